// Tile binning: exclusive scan of per-Gaussian tile counts, (tile|depth) key emission and
// tile-range extraction.  No counterpart in the reference (its CUDA path scatters with
// atomics, src/renderer.cu:89-103); the contract is oracle/bins_oracle.c, bit for bit.
#include "common.cuh"

namespace b2s {

// Level 2 of the scan: one block turns the per-preprocess-block sums into exclusive
// offsets in place, and publishes the total / overflow counters.
__global__ void __launch_bounds__(1024)
scan_bsum_kernel(long long* __restrict__ bsum, int nb, long long max_pairs, Counters* __restrict__ counters) {
  __shared__ long long wtot[32];
  __shared__ long long carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base = 0; base < nb; base += 1024) {
    const int i = base + threadIdx.x;
    const long long v = (i < nb) ? bsum[i] : 0;
    long long x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wtot[wid] = x;
    __syncthreads();
    if (wid == 0) {
      long long t = wtot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      wtot[lane] = t;  // inclusive over warps
    }
    __syncthreads();
    const long long carry = carry_s;
    const long long excl = carry + (wid > 0 ? wtot[wid - 1] : 0) + (x - v);
    if (i < nb) bsum[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + wtot[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const long long total = carry_s;
    counters->needed = total;
    counters->kept = (total <= max_pairs) ? (int)total : 0;   // on overflow nothing is rendered; the caller must look at the flag
    counters->overflow = total > max_pairs ? 1 : 0;
  }
}

// Key emission.  Block b re-scans the counts of its PRE_BLOCK Gaussians, adds the block
// offset, and each thread writes its Gaussian's tiles row-major:
//   key = tile_id << 32 | depth_bits ,  val = Gaussian index.
// A Gaussian whose slots would cross max_pairs is dropped whole (overflow was flagged).
__global__ void __launch_bounds__(PRE_BLOCK)
emit_kernel(int n, int tiles_x, long long max_pairs, const uint2* __restrict__ rect,
            const uint32_t* __restrict__ dbits, const int* __restrict__ cnt,
            const long long* __restrict__ bsum, unsigned long long* __restrict__ keys, int* __restrict__ vals) {
  __shared__ int wtot[PRE_BLOCK / 32];
  const int i = blockIdx.x * PRE_BLOCK + threadIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int c = (i < n) ? cnt[i] : 0;
  int x = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) wtot[wid] = x;
  __syncthreads();
  int wbase = 0;
#pragma unroll
  for (int q = 0; q < PRE_BLOCK / 32; ++q) wbase += (q < wid) ? wtot[q] : 0;
  if (c == 0) return;
  long long o = bsum[blockIdx.x] + wbase + (x - c);
  if (o + c > max_pairs) return;
  const uint2 rc = rect[i];
  const int tx0 = rc.x & 0xffff, ty0 = rc.x >> 16, tx1 = rc.y & 0xffff, ty1 = rc.y >> 16;
  const unsigned long long d = dbits[i];
  for (int ty = ty0; ty <= ty1; ++ty)
    for (int tx = tx0; tx <= tx1; ++tx) {
      keys[o] = ((unsigned long long)(uint32_t)(ty * tiles_x + tx) << 32) | d;
      vals[o] = i;
      ++o;
    }
}

int launch_bin(const ViewParams& vp, int n, int64_t max_pairs, const uint2* rect, const uint32_t* dbits,
               const int* cnt, long long* bsum, unsigned long long* keys, int* vals, Counters* counters,
               cudaStream_t st) {
  const int nb = (n + PRE_BLOCK - 1) / PRE_BLOCK;
  scan_bsum_kernel<<<1, 1024, 0, st>>>(bsum, nb, (long long)max_pairs, counters);
  B2S_LAUNCH_CHECK();
  if (n > 0 && keys != nullptr) {
    emit_kernel<<<nb, PRE_BLOCK, 0, st>>>(n, vp.tiles_x, (long long)max_pairs, rect, dbits, cnt, bsum, keys, vals);
    B2S_LAUNCH_CHECK();
  }
  return B2S_OK;
}

// Work-unit table: tile t gets max(1, ceil(count_t / SEG)) units (an empty tile keeps one so
// that its background pixels are written).  unit_start[t] = exclusive prefix, unit_start[n_tiles]
// = number of units; units[u] = (tile, segment).  One block: n_tiles is a few thousand.
__global__ void __launch_bounds__(1024)
units_kernel(const int2* __restrict__ ranges, int n_tiles, int unit_cap, int* __restrict__ unit_start,
             int2* __restrict__ units) {
  __shared__ int wtot[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base = 0; base < n_tiles; base += 1024) {
    const int t = base + threadIdx.x;
    int v = 0;
    if (t < n_tiles) {
      const int2 rg = ranges[t];
      const int c = rg.y - rg.x;
      v = c > 0 ? (c + SEG - 1) / SEG : 1;
    }
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wtot[wid] = x;
    __syncthreads();
    if (wid == 0) {
      int s = wtot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += y;
      }
      wtot[lane] = s;
    }
    __syncthreads();
    const int carry = carry_s;
    const int excl = carry + (wid > 0 ? wtot[wid - 1] : 0) + (x - v);
    if (t < n_tiles) {
      unit_start[t] = excl;
      for (int s = 0; s < v; ++s)
        if (excl + s < unit_cap) units[excl + s] = make_int2(t, s);
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + wtot[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) unit_start[n_tiles] = carry_s < unit_cap ? carry_s : unit_cap;
}

int launch_units(const int2* ranges, int n_tiles, int64_t unit_cap, int* unit_start, int2* units, cudaStream_t st) {
  units_kernel<<<1, 1024, 0, st>>>(ranges, n_tiles, (int)unit_cap, unit_start, units);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

// ranges[t] = [start, end) of tile t in the sorted list; (0,0) for empty tiles (memset before).
__global__ void ranges_kernel(const unsigned long long* __restrict__ keys, const int* __restrict__ count_dev,
                              int2* __restrict__ ranges) {
  const int m = *count_dev;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const uint32_t t = (uint32_t)(keys[i] >> 32);
  if (i == 0 || (uint32_t)(keys[i - 1] >> 32) != t) ranges[t].x = i;
  if (i == m - 1 || (uint32_t)(keys[i + 1] >> 32) != t) ranges[t].y = i + 1;
}

int launch_ranges(const unsigned long long* keys, const int* count_dev, int64_t cap, int n_tiles, int2* ranges,
                  cudaStream_t st) {
  B2S_CUDA_TRY(cudaMemsetAsync(ranges, 0, (size_t)n_tiles * sizeof(int2), st));
  if (cap <= 0) return B2S_OK;
  const int threads = 256;
  const int blocks = (int)((cap + threads - 1) / threads);
  ranges_kernel<<<blocks, threads, 0, st>>>(keys, count_dev, ranges);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

}  // namespace b2s
