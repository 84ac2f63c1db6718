// Stable LSD radix sort of (uint64 key, int32 value) pairs, 8 bits per pass.
// Per pass: per-block digit histogram -> exclusive scan over (digit, block) -> ranked scatter.
// The element count lives on the device (no host sync); grids are sized by capacity and
// blocks past the live count exit.  HBM-bound: per pass 2 key reads + 1 value read + 1 pair
// write = 32 B per pair.
#include "common.cuh"

namespace b2s {

constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = SORT_KPB / SORT_THREADS;   // 16
constexpr int SORT_WARPS = SORT_THREADS / 32;         // 8
constexpr int SCAN_ELEMS = 4096;                      // elements per scan block

__global__ void __launch_bounds__(SORT_THREADS)
radix_hist_kernel(const unsigned long long* __restrict__ keys, const int* __restrict__ count_dev, int shift,
                  unsigned mask, int nb, int* __restrict__ counts) {
  __shared__ int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int m = *count_dev;
  const long long base = (long long)blockIdx.x * SORT_KPB;
#pragma unroll 4
  for (int j = 0; j < SORT_ITEMS; ++j) {
    const long long idx = base + j * SORT_THREADS + threadIdx.x;
    const bool valid = idx < m;
    const unsigned d = valid ? (unsigned)((keys[valid ? idx : 0] >> shift) & mask) : 256u;
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    if (valid && (__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&h[d], __popc(peers));
  }
  __syncthreads();
  counts[threadIdx.x * nb + blockIdx.x] = h[threadIdx.x];
}

// ---- generic exclusive scan of int32 (3 kernels) ----------------------------------------------
__global__ void __launch_bounds__(256) scan_reduce_kernel(const int* __restrict__ data, int len, int* __restrict__ bs) {
  __shared__ int ws[8];
  const int base = blockIdx.x * SCAN_ELEMS;
  int s = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ELEMS / 256; ++j) {
    const int i = base + j * 256 + threadIdx.x;
    s += (i < len) ? data[i] : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += ws[q];
    bs[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(1024) scan_small_kernel(int* __restrict__ bs, int nb) {
  __shared__ int wtot[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base = 0; base < nb; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = (i < nb) ? bs[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wtot[wid] = x;
    __syncthreads();
    if (wid == 0) {
      int t = wtot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      wtot[lane] = t;
    }
    __syncthreads();
    const int carry = carry_s;
    if (i < nb) bs[i] = carry + (wid > 0 ? wtot[wid - 1] : 0) + (x - v);
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + wtot[31];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) scan_apply_kernel(int* __restrict__ data, int len, const int* __restrict__ bs) {
  __shared__ int ws[8];
  constexpr int PER = SCAN_ELEMS / 256;   // 16 consecutive elements per thread
  const int base = blockIdx.x * SCAN_ELEMS + threadIdx.x * PER;
  int v[PER];
  int s = 0;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    v[j] = (base + j < len) ? data[base + j] : 0;
    s += v[j];
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int x = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) ws[wid] = x;
  __syncthreads();
  int run = bs[blockIdx.x] + (x - s);
#pragma unroll
  for (int q = 0; q < 8; ++q) run += (q < wid) ? ws[q] : 0;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    if (base + j < len) data[base + j] = run;
    run += v[j];
  }
}

static int scan_i32_inplace(int* data, int len, int* bs, cudaStream_t st) {
  const int nb = (len + SCAN_ELEMS - 1) / SCAN_ELEMS;
  scan_reduce_kernel<<<nb, 256, 0, st>>>(data, len, bs);
  B2S_LAUNCH_CHECK();
  scan_small_kernel<<<1, 1024, 0, st>>>(bs, nb);
  B2S_LAUNCH_CHECK();
  scan_apply_kernel<<<nb, 256, 0, st>>>(data, len, bs);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

// Ranked scatter.  Warp w of a block owns 512 consecutive keys, item j of lane l is key
// w*512 + j*32 + l, so (item, lane) order == index order and the ranks below are stable.
__global__ void __launch_bounds__(SORT_THREADS)
radix_scatter_kernel(const unsigned long long* __restrict__ keys_in, const int* __restrict__ vals_in,
                     unsigned long long* __restrict__ keys_out, int* __restrict__ vals_out,
                     const int* __restrict__ count_dev, int shift, unsigned mask, int nb,
                     const int* __restrict__ offsets) {
  __shared__ int wh[SORT_WARPS][256];
  const int m = *count_dev;
  const long long base = (long long)blockIdx.x * SORT_KPB;
  if (base >= m) return;
  for (int q = threadIdx.x; q < SORT_WARPS * 256; q += SORT_THREADS) (&wh[0][0])[q] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const long long wbase = base + (long long)w * (32 * SORT_ITEMS);
  unsigned long long k[SORT_ITEMS];
  int rank[SORT_ITEMS];
#pragma unroll
  for (int j = 0; j < SORT_ITEMS; ++j) {
    const long long idx = wbase + j * 32 + lane;
    const bool valid = idx < m;
    k[j] = valid ? keys_in[idx] : ~0ull;
    const unsigned d = valid ? (unsigned)((k[j] >> shift) & mask) : 256u;
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int r = __popc(peers & lt);
    const int pre = valid ? wh[w][d] : 0;
    __syncwarp();
    if (valid && r == 0) wh[w][d] = pre + __popc(peers);
    __syncwarp();
    rank[j] = pre + r;
  }
  __syncthreads();
  {
    const int d = threadIdx.x;
    int run = offsets[d * nb + blockIdx.x];
#pragma unroll
    for (int q = 0; q < SORT_WARPS; ++q) {
      const int t = wh[q][d];
      wh[q][d] = run;
      run += t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < SORT_ITEMS; ++j) {
    const long long idx = wbase + j * 32 + lane;
    if (idx < m) {
      const unsigned d = (unsigned)((k[j] >> shift) & mask);
      const int pos = wh[w][d] + rank[j];
      keys_out[pos] = k[j];
      vals_out[pos] = vals_in[idx];
    }
  }
}

// exclusive scan of `len` ints in place; bs: (len/4096 + 2) ints of scratch
int launch_scan_i32(int* data, int len, int* bs, cudaStream_t st) { return len > 0 ? scan_i32_inplace(data, len, bs, st) : B2S_OK; }

int launch_sort(unsigned long long* keysA, int* valsA, unsigned long long* keysB, int* valsB, int64_t cap,
                const int* count_dev, int begin_bit, int end_bit, int* hist, int* hsum, int* result_in_B,
                cudaStream_t st) {
  *result_in_B = 0;
  if (cap <= 0) return B2S_OK;
  const int nb = (int)((cap + SORT_KPB - 1) / SORT_KPB);
  unsigned long long* ks = keysA;
  int* vs = valsA;
  unsigned long long* kd = keysB;
  int* vd = valsB;
  for (int bit = begin_bit; bit < end_bit; bit += 8) {
    const int nbits = (end_bit - bit) < 8 ? (end_bit - bit) : 8;
    const unsigned mask = (1u << nbits) - 1u;
    radix_hist_kernel<<<nb, SORT_THREADS, 0, st>>>(ks, count_dev, bit, mask, nb, hist);
    B2S_LAUNCH_CHECK();
    const int rc = scan_i32_inplace(hist, nb * 256, hsum, st);
    if (rc != B2S_OK) return rc;
    radix_scatter_kernel<<<nb, SORT_THREADS, 0, st>>>(ks, vs, kd, vd, count_dev, bit, mask, nb, hist);
    B2S_LAUNCH_CHECK();
    unsigned long long* tk = ks; ks = kd; kd = tk;
    int* tv = vs; vs = vd; vd = tv;
    *result_in_B ^= 1;
  }
  return B2S_OK;
}

}  // namespace b2s
