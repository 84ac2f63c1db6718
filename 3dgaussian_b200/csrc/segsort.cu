// Depth order inside the tiles (MODE_SORTED) without a multi-pass sort of the (Gaussian, tile) pairs.
//
// The reference orders ALL Gaussians by camera z once (std::sort, src/renderer_cpu.cpp:131-146) and every pixel walks
// that order.  The same idea on the tile lists, in three steps:
//   1. the N Gaussians are partitioned into nb DEPTH SLABS (nb = the counting sort's block count): splitters from a
//      sorted sample of 4096 depth words, slab(i) = number of splitters below depth_bits(i) -- monotone in depth, equal
//      depths share a slab, sizes balanced by the sample; one count and one scatter pass over N (no order inside a slab);
//   2. the tile-major counting sort of bin.cu walks the Gaussians slab by slab: block b of the histogram / scatter
//      kernels owns slab b, so inside a tile's list the sub-range written by block b ("group" (b, tile), delimited by
//      the prefix table the counting sort already has) precedes block b+1's in depth;
//   3. only the inside of a group is unordered (the scatter's shared-memory atomics): group_sort_kernel sorts every
//      group by the composite key depth_bits << 32 | id -- a warp per group (a few dozen elements), the CTA for the rare
//      group that outgrows a warp's slab.
// The result is exactly the order of a stable LSD radix sort of the full 64-bit (tile | depth) keys emitted in Gaussian
// order, i.e. oracle/bins_oracle.c's, bit for bit (tests/test_gpu_parity.py::test_bins_bit_exact_*).
#include "common.cuh"

namespace b2s {

constexpr int GS_THREADS = 256;
constexpr int GS_WARPS = GS_THREADS / 32;
constexpr int GS_WARP_CAP = 512;                   // elements a warp sorts in its own slab (4 KB)
constexpr int GS_CTA_CAP = GS_WARPS * GS_WARP_CAP;   // elements the CTA sorts at a time (all slabs: 32 KB)
constexpr int GS_MAX_BIG = 64;                     // oversized groups remembered per tile before the CTA handles them
constexpr int GS_SPLIT = 16;                       // CTAs per tile (grid.y): each takes a contiguous share of the blocks
constexpr int GS_MAX_SHARE = (CS_NB + GS_SPLIT - 1) / GS_SPLIT;   // groups per CTA

__device__ __forceinline__ int next_pow2(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

// bitonic sort of P (power of two) keys in shared memory by NT cooperating threads (NT = 32: one warp, __syncwarp;
// NT = GS_THREADS: the CTA, __syncthreads)
template <int NT>
__device__ __forceinline__ void bitonic_sort(unsigned long long* s, int P, int t) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = t; i < (P >> 1); i += NT) {
        const int lo = 2 * i - (i & (j - 1));          // bit j clear
        const int hi = lo + j;
        const unsigned long long a = s[lo], b = s[hi];
        const bool up = (lo & k) == 0;
        if ((a > b) == up) { s[lo] = b; s[hi] = a; }
      }
      if (NT == 32) __syncwarp(); else __syncthreads();
    }
  }
}

// Bitonic sort of up to 32 R keys held in registers by one warp: element r * 32 + lane lives in k[r] of `lane`.
// Stages with a partner distance below 32 exchange through shuffles, the others between the thread's own registers:
// no shared memory, no barriers -- ~8 instructions per key per stage, against ~20 for the shared-memory version, which
// matters because the (block, tile) groups are small (a few dozen elements) and there are > 100 000 of them per frame.
template <int R>
__device__ __forceinline__ void warp_sort_regs(unsigned long long (&k)[R], int lane) {
#pragma unroll
  for (int kk = 2; kk <= 32 * R; kk <<= 1) {
#pragma unroll
    for (int j = kk >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        const int dr = j >> 5;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if ((r & dr) == 0) {
            const bool up = (((r << 5) | lane) & kk) == 0;
            const unsigned long long a = k[r], b = k[r | dr];
            if ((a > b) == up) { k[r] = b; k[r | dr] = a; }
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const unsigned long long mine = k[r];
          const unsigned long long other = __shfl_xor_sync(0xffffffffu, mine, j);
          const bool up = (((r << 5) | lane) & kk) == 0;
          const bool lower = (lane & j) == 0;
          const bool take_min = lower == up;
          k[r] = (mine < other) == take_min ? mine : other;
        }
      }
    }
  }
}

template <int R>
__device__ __forceinline__ void warp_sort_group(int* __restrict__ dst, int m, int lane, const uint32_t* __restrict__ dbits) {
  unsigned long long k[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int i = r * 32 + lane;
    k[r] = ~0ull;
    if (i < m) {
      const int id = dst[i];
      k[r] = ((unsigned long long)__ldg(dbits + id) << 32) | (unsigned)id;
    }
  }
  warp_sort_regs<R>(k, lane);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int i = r * 32 + lane;
    if (i < m) dst[i] = (int)(unsigned)(k[r] & 0xffffffffull);
  }
}

template <int NT>
__device__ __forceinline__ void load_keys(unsigned long long* s, int P, int t, const int* __restrict__ src, int len,
                                          const uint32_t* __restrict__ dbits) {
  for (int i = t; i < P; i += NT) {
    unsigned long long k = ~0ull;
    if (i < len) {
      const int id = src[i];
      k = ((unsigned long long)__ldg(dbits + id) << 32) | (unsigned)id;
    }
    s[i] = k;
  }
  if (NT == 32) __syncwarp(); else __syncthreads();
}

// One CTA per (tile, share of the blocks).  The share's groups are contiguous in the tile's list, so the CTA stages all
// of them at once -- ids with coalesced loads, depth words with one gather, the group boundaries from the prefix table --
// and the memory latency is paid twice per CTA instead of four times per group; the warps then sort the groups out of
// shared memory (registers + shuffles up to 128 elements, the warp's slab up to 512, the CTA beyond) and the result
// goes back with coalesced stores.
__global__ void __launch_bounds__(GS_THREADS)
group_sort_kernel(const int* __restrict__ table, const int* __restrict__ total, const int2* __restrict__ ranges,
                  const int* __restrict__ ne_list, int nb,
                  int n_tiles, const uint32_t* __restrict__ dbits, const Counters* __restrict__ counters,
                  int* __restrict__ vals, unsigned long long* __restrict__ scratch) {
  __shared__ unsigned long long keys[GS_CTA_CAP];          // 32 KB: the share's composite keys (or the CTA-wide sort buffer)
  __shared__ int bound[GS_MAX_SHARE + 1];
  __shared__ int2 big[GS_MAX_BIG];
  __shared__ int nbig_s;
  if (counters->overflow) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int share = (nb + GS_SPLIT - 1) / GS_SPLIT;        // <= GS_MAX_SHARE (nb <= CS_NB)
  // persistent grid over the compact list of (non-empty tile, share) items written by the tile scan: a CTA per
  // (tile, share) of the whole grid spent more time retiring the empty tiles' CTAs than sorting
  const int n_items = ne_list[0] * GS_SPLIT;
  for (int w = blockIdx.x; w < n_items; w += gridDim.x) {     // block uniform
  __syncthreads();                                          // the previous item's shared state is no longer read
  const int tile = ne_list[1 + w / GS_SPLIT];
  const int2 rg = ranges[tile];
  const int b_lo = min(nb, (w % GS_SPLIT) * share), b_hi = min(nb, b_lo + share);
  const int ng = b_hi - b_lo;
  if (ng <= 0) continue;
  if (threadIdx.x <= ng) {
    const int b = b_lo + threadIdx.x;
    bound[threadIdx.x] = (b < nb) ? table[(size_t)b * n_tiles + tile] : total[tile];
  }
  if (threadIdx.x == 0) nbig_s = 0;
  __syncthreads();
  const int s_lo = bound[0], m_all = bound[ng] - s_lo;
  if (m_all <= 1) continue;
  int* base = vals + rg.x + s_lo;
  const bool staged = m_all <= GS_CTA_CAP;
  if (staged) {
    for (int i = threadIdx.x; i < m_all; i += GS_THREADS) {
      const int id = base[i];
      keys[i] = ((unsigned long long)__ldg(dbits + id) << 32) | (unsigned)id;
    }
    __syncthreads();
  }
  // ---- phase 1: a warp per group
  for (int g = warp; g < ng; g += GS_WARPS) {
    const int o = bound[g] - s_lo, m = bound[g + 1] - bound[g];
    if (m <= 1) continue;
    if (staged && m <= 128) {
      unsigned long long k[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) k[r] = (r * 32 + lane < m) ? keys[o + r * 32 + lane] : ~0ull;
      if (m <= 32) { unsigned long long k1[1] = {k[0]}; warp_sort_regs<1>(k1, lane); k[0] = k1[0]; }
      else if (m <= 64) { unsigned long long k2[2] = {k[0], k[1]}; warp_sort_regs<2>(k2, lane); k[0] = k2[0]; k[1] = k2[1]; }
      else warp_sort_regs<4>(k, lane);
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (r * 32 + lane < m) keys[o + r * 32 + lane] = k[r];
      continue;
    }
    if (lane == 0) {                                         // larger (or unstaged) groups: the CTA takes them below
      const int q = atomicAdd(&nbig_s, 1);
      if (q < GS_MAX_BIG) big[q] = make_int2(bound[g] , m);
    }
  }
  __syncthreads();
  if (staged) {
    for (int i = threadIdx.x; i < m_all; i += GS_THREADS) base[i] = (int)(unsigned)(keys[i] & 0xffffffffull);
    __syncthreads();
  }
  const int nbig = nbig_s;
  if (nbig == 0) continue;
  // ---- phase 2: the CTA sorts the oversized groups one after the other straight from global memory (the staged keys
  // are no longer needed).  More of them than the list holds: the CTA rescans its groups.
  unsigned long long* all = keys;
  const int rounds = nbig <= GS_MAX_BIG ? nbig : ng;
  for (int rr = 0; rr < rounds; ++rr) {
    int s0, m;
    if (nbig <= GS_MAX_BIG) { s0 = big[rr].x; m = big[rr].y; }
    else {
      s0 = bound[rr];
      m = bound[rr + 1] - bound[rr];
      if (m <= 1 || (staged && m <= 128)) continue;         // block uniform
    }
    int* dst = vals + rg.x + s0;
    if (m <= GS_CTA_CAP) {
      const int P = next_pow2(m);
      load_keys<GS_THREADS>(all, P, threadIdx.x, dst, m, dbits);
      bitonic_sort<GS_THREADS>(all, P, threadIdx.x);
      for (int i = threadIdx.x; i < m; i += GS_THREADS) dst[i] = (int)(unsigned)(all[i] & 0xffffffffull);
      __syncthreads();
      continue;
    }
    // longer than one shared-memory sort: sorted chunks into the scratch array, then placement by rank (own index in
    // its chunk + lower bounds in the other chunks; the keys are unique)
    const int C = (m + GS_CTA_CAP - 1) / GS_CTA_CAP;
    unsigned long long* out = scratch + rg.x + s0;
    for (int c = 0; c < C; ++c) {
      const int cb = c * GS_CTA_CAP, len = min(GS_CTA_CAP, m - cb);
      const int P = next_pow2(len);
      load_keys<GS_THREADS>(all, P, threadIdx.x, dst + cb, len, dbits);
      bitonic_sort<GS_THREADS>(all, P, threadIdx.x);
      for (int i = threadIdx.x; i < len; i += GS_THREADS) out[cb + i] = all[i];
      __syncthreads();
    }
    __threadfence_block();
    __syncthreads();
    for (int idx = threadIdx.x; idx < m; idx += GS_THREADS) {
      const unsigned long long k = out[idx];
      const int a = idx / GS_CTA_CAP;
      int pos = idx - a * GS_CTA_CAP;
      for (int c = 0; c < C; ++c) {
        if (c == a) continue;
        const unsigned long long* ch = out + c * GS_CTA_CAP;
        int lo = 0, hi = min(GS_CTA_CAP, m - c * GS_CTA_CAP);
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (ch[mid] < k) lo = mid + 1; else hi = mid;
        }
        pos += lo;
      }
      dst[pos] = (int)(unsigned)(k & 0xffffffffull);
    }
    __syncthreads();
  }
  }   // work items
}

// ---- step 1: depth slabs -----------------------------------------------------------------------------------------
constexpr int SL_SAMPLES = 2048;
constexpr int SL_MAXB = 512;          // >= CS_NB

// one CTA: sorted sample -> nb-1 splitters; clears the slab counters
__global__ void __launch_bounds__(1024)
slab_sample_kernel(const uint32_t* __restrict__ dbits, int n, int nb, uint32_t* __restrict__ splitters, int* __restrict__ count) {
  __shared__ uint32_t smp[SL_SAMPLES];
  for (int j = threadIdx.x; j < SL_SAMPLES; j += 1024) smp[j] = dbits[(int)(((long long)j * n) / SL_SAMPLES)];
  for (int b = threadIdx.x; b <= nb; b += 1024) count[b] = 0;
  __syncthreads();
  for (int k = 2; k <= SL_SAMPLES; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < SL_SAMPLES / 2; i += 1024) {
        const int lo = 2 * i - (i & (j - 1)), hi = lo + j;
        const uint32_t a = smp[lo], c = smp[hi];
        if ((a > c) == ((lo & k) == 0)) { smp[lo] = c; smp[hi] = a; }
      }
      __syncthreads();
    }
  for (int b = threadIdx.x; b < nb - 1; b += 1024) splitters[b] = smp[(int)(((long long)(b + 1) * SL_SAMPLES) / nb)];
}

__device__ __forceinline__ int slab_of(const uint32_t* spl, int nspl, uint32_t key) {   // number of splitters < key
  int lo = 0, hi = nspl;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (spl[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// slab of every Gaussian (kept in slab_id) + slab sizes
__global__ void __launch_bounds__(256)
slab_count_kernel(const uint32_t* __restrict__ dbits, int n, int nb, const uint32_t* __restrict__ splitters,
                  int* __restrict__ slab_id, int* __restrict__ count) {
  __shared__ uint32_t spl[SL_MAXB];
  __shared__ int hist[SL_MAXB];
  for (int b = threadIdx.x; b < nb; b += 256) { hist[b] = 0; if (b < nb - 1) spl[b] = splitters[b]; }
  __syncthreads();
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const int sl = slab_of(spl, nb - 1, dbits[i]);
    slab_id[i] = sl;
    atomicAdd(&hist[sl], 1);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < nb; b += 256)
    if (hist[b] != 0) atomicAdd(&count[b], hist[b]);
}

// one CTA: slab_start[0..nb] = exclusive prefix of the sizes; cursor[b] = slab_start[b]
__global__ void __launch_bounds__(SL_MAXB)
slab_scan_kernel(const int* __restrict__ count, int nb, int* __restrict__ slab_start, int* __restrict__ cursor) {
  __shared__ int sc[SL_MAXB];
  const int t = threadIdx.x;
  sc[t] = t < nb ? count[t] : 0;
  __syncthreads();
  for (int o = 1; o < SL_MAXB; o <<= 1) {
    const int v = t >= o ? sc[t - o] : 0;
    __syncthreads();
    sc[t] += v;
    __syncthreads();
  }
  const int excl = t > 0 ? sc[t - 1] : 0;
  if (t <= nb) { slab_start[t] = excl; if (t < nb) cursor[t] = excl; }
}

// order[slab_start[s] ..) = the Gaussians of slab s (any order): a block reserves its share of every slab with one
// global atomic per (block, slab) and ranks its elements with shared-memory atomics
__global__ void __launch_bounds__(256)
slab_scatter_kernel(const int* __restrict__ slab_id, int n, int nb, int per_block, int* __restrict__ cursor, int* __restrict__ order) {
  __shared__ int hist[SL_MAXB];
  const int i0 = blockIdx.x * per_block, i1 = min(n, i0 + per_block);
  for (int b = threadIdx.x; b < nb; b += 256) hist[b] = 0;
  __syncthreads();
  for (int i = i0 + threadIdx.x; i < i1; i += 256) atomicAdd(&hist[slab_id[i]], 1);
  __syncthreads();
  for (int b = threadIdx.x; b < nb; b += 256) {
    const int c = hist[b];
    hist[b] = c != 0 ? atomicAdd(&cursor[b], c) : 0;
  }
  __syncthreads();
  for (int i = i0 + threadIdx.x; i < i1; i += 256) order[atomicAdd(&hist[slab_id[i]], 1)] = i;
}

// keys[pos] = tile << 32 | depth bits of the Gaussian at pos (b2s_dump_bins: the sorted keys of the bit-exact tests)
__global__ void __launch_bounds__(256)
rebuild_keys_kernel(const int2* __restrict__ ranges, const uint32_t* __restrict__ dbits, const int* __restrict__ vals,
                    unsigned long long* __restrict__ keys) {
  const int2 rg = ranges[blockIdx.x];
  for (int i = rg.x + threadIdx.x; i < rg.y; i += 256)
    keys[i] = ((unsigned long long)blockIdx.x << 32) | __ldg(dbits + vals[i]);
}

// Step 1: order[] = the Gaussians slab by slab, slab_start[0..nb] = the slab boundaries (device).
int launch_depth_slabs(const uint32_t* dbits, int n, int nb, uint32_t* splitters, int* count, int* slab_start, int* cursor,
                       int* slab_id, int* order, cudaStream_t st) {
  if (n <= 0) return B2S_OK;
  if (nb > SL_MAXB) { set_error("depth slabs: %d blocks exceed %d", nb, SL_MAXB); return B2S_ERR_INVALID; }
  slab_sample_kernel<<<1, 1024, 0, st>>>(dbits, n, nb, splitters, count);
  B2S_LAUNCH_CHECK();
  int blocks = (n + 255) / 256;
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  slab_count_kernel<<<blocks, 256, 0, st>>>(dbits, n, nb, splitters, slab_id, count);
  B2S_LAUNCH_CHECK();
  slab_scan_kernel<<<1, SL_MAXB, 0, st>>>(count, nb, slab_start, cursor);
  B2S_LAUNCH_CHECK();
  const int per_block = 4096;
  slab_scatter_kernel<<<(n + per_block - 1) / per_block, 256, 0, st>>>(slab_id, n, nb, per_block, cursor, order);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

// Step 3 (after the order-indirected counting sort): sort the inside of every (block, tile) group.
int launch_group_sort(const ViewParams& vp, const int* table, const int* total, const int2* ranges, const int* ne_list, int nb,
                      const uint32_t* dbits, const Counters* counters, int* vals, unsigned long long* scratch,
                      unsigned long long* keys_out, cudaStream_t st) {
  if (vp.n_tiles <= 0) return B2S_OK;
  group_sort_kernel<<<sm_count() * 7, GS_THREADS, 0, st>>>(table, total, ranges, ne_list, nb, vp.n_tiles, dbits, counters, vals, scratch);
  B2S_LAUNCH_CHECK();
  if (keys_out != nullptr) {
    rebuild_keys_kernel<<<vp.n_tiles, 256, 0, st>>>(ranges, dbits, vals, keys_out);
    B2S_LAUNCH_CHECK();
  }
  return B2S_OK;
}

}  // namespace b2s
