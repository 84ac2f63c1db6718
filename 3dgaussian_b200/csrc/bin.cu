// Tile binning: exclusive scan of per-Gaussian tile counts, (tile|depth) key emission and
// tile-range extraction.  No counterpart in the reference (its CUDA path scatters with
// atomics, src/renderer.cu:89-103); the contract is oracle/bins_oracle.c, bit for bit.
#include <stdlib.h>

#include "common.cuh"

namespace b2s {

// unit descriptor table (common.cuh: launch_udesc): a unit's class is its size in eighths of the largest unit
constexpr int UD_NCLS = 8;

// Visits the tiles of Gaussian i row-major: the set bits of its mask for rects of at most 8 x 8 tiles
// (tile_cull_mask), the whole rect otherwise.
template <typename F>
__device__ __forceinline__ void for_each_tile(uint2 rc, unsigned long long mask, int tiles_x, F f) {
  const int tx0 = rc.x & 0xffff, ty0 = rc.x >> 16, tx1 = rc.y & 0xffff, ty1 = rc.y >> 16;
  const int w = tx1 - tx0 + 1, h = ty1 - ty0 + 1;
  if (w <= 0 || h <= 0) return;
  if (w <= 8 && h <= 8) {
    const unsigned row_bits = (1u << w) - 1u;
    for (int r = 0; r < h; ++r) {
      unsigned rb = (unsigned)(mask >> (r * w)) & row_bits;
      const int base = (ty0 + r) * tiles_x + tx0;
      while (rb) {
        const int c = __ffs((int)rb) - 1;
        rb &= rb - 1;
        f(base + c);
      }
    }
  } else {
    for (int ty = ty0; ty <= ty1; ++ty)
      for (int tx = tx0; tx <= tx1; ++tx) f(ty * tiles_x + tx);
  }
}

// Level 2 of the scan: one block turns the per-preprocess-block sums into exclusive
// offsets in place, and publishes the total / overflow counters.
__global__ void __launch_bounds__(1024)
scan_bsum_kernel(long long* __restrict__ bsum, int nb, long long max_pairs, Counters* __restrict__ counters,
                 Counters* __restrict__ mirror) {
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
    if (mirror != nullptr) {
      mirror->needed = total;
      mirror->kept = (total <= max_pairs) ? (int)total : 0;
      mirror->overflow = total > max_pairs ? 1 : 0;
    }
  }
}

// Key emission.  Block b re-scans the counts of its PRE_BLOCK Gaussians, adds the block
// offset, and each thread writes its Gaussian's tiles row-major:
//   key = tile_id << 32 | depth_bits ,  val = Gaussian index.
// A Gaussian whose slots would cross max_pairs is dropped whole (overflow was flagged).
__global__ void __launch_bounds__(PRE_BLOCK)
emit_kernel(int n, int tiles_x, long long max_pairs, const uint2* __restrict__ rect,
            const unsigned long long* __restrict__ tmask, const uint32_t* __restrict__ dbits, const int* __restrict__ cnt,
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
  const unsigned long long d = dbits[i];
  for_each_tile(rect[i], tmask[i], tiles_x, [&](int tile) {
    keys[o] = ((unsigned long long)(uint32_t)tile << 32) | d;
    vals[o] = i;
    ++o;
  });
}

int launch_bin(const ViewParams& vp, int n, int64_t max_pairs, const uint2* rect, const unsigned long long* tmask,
               const uint32_t* dbits,
               const int* cnt, long long* bsum, unsigned long long* keys, int* vals, Counters* counters,
               Counters* mirror, cudaStream_t st) {
  const int nb = (n + PRE_BLOCK - 1) / PRE_BLOCK;
  scan_bsum_kernel<<<1, 1024, 0, st>>>(bsum, nb, (long long)max_pairs, counters, mirror);
  B2S_LAUNCH_CHECK();
  if (n > 0 && keys != nullptr) {
    emit_kernel<<<nb, PRE_BLOCK, 0, st>>>(n, vp.tiles_x, (long long)max_pairs, rect, tmask, dbits, cnt, bsum, keys, vals);
    B2S_LAUNCH_CHECK();
  }
  return B2S_OK;
}

// Work-unit table: tile t gets max(1, ceil(count_t / SEG)) units (an empty tile keeps one so
// that its background pixels are written).  unit_start[t] = exclusive prefix, unit_start[n_tiles]
// = number of units; units[u] = (tile, segment).  One block: n_tiles is a few thousand.
__global__ void __launch_bounds__(1024)
units_kernel(const int2* __restrict__ ranges, int n_tiles, int SEG, int unit_cap, int* __restrict__ unit_start,
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

int launch_units(const int2* ranges, int n_tiles, int seg, int64_t unit_cap, int* unit_start, int2* units, cudaStream_t st) {
  units_kernel<<<1, 1024, 0, st>>>(ranges, n_tiles, seg, (int)unit_cap, unit_start, units);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

// ---- unit descriptor table (see common.cuh: launch_udesc) ---------------------------------------------------
// One block.  A unit's class is its number of 128-Gaussian steps (1 .. SEG/128); thread t owns the contiguous
// chunk of tiles [t*per, (t+1)*per), counts its units per class, one block-wide exclusive scan per class gives it
// a deterministic slot range inside each class, classes are laid out largest first.
__global__ void __launch_bounds__(1024)
udesc_kernel(const int2* __restrict__ ranges, const int* __restrict__ unit_start, int n_tiles, int SEG, int unit_cap,
             int4* __restrict__ udesc, Counters* __restrict__ counters) {
  const int UD_STEP = (SEG + UD_NCLS - 1) / UD_NCLS;
  __shared__ int wtot[UD_NCLS][32];
  __shared__ int cls_total[UD_NCLS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int per = (n_tiles + 1023) / 1024;
  const int t0 = min(n_tiles, (int)threadIdx.x * per), t1 = min(n_tiles, t0 + per);
  int cnt[UD_NCLS];
#pragma unroll
  for (int k = 0; k < UD_NCLS; ++k) cnt[k] = 0;
  for (int t = t0; t < t1; ++t) {
    const int2 rg = ranges[t];
    const int c = rg.y - rg.x;
    if (c <= 0) continue;
    const int v = (c + SEG - 1) / SEG;
    if (unit_start[t] + v > unit_cap) continue;            // cannot happen with max_units(); keeps the table in bounds
    cnt[UD_NCLS - 1] += v - 1;                              // full units
    const int last = c - (v - 1) * SEG;
    const int kl = (last + UD_STEP - 1) / UD_STEP - 1;
#pragma unroll
    for (int k = 0; k < UD_NCLS; ++k) cnt[k] += (k == kl) ? 1 : 0;
  }
  int excl[UD_NCLS];
#pragma unroll
  for (int k = 0; k < UD_NCLS; ++k) {
    int x = cnt[k];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wtot[k][wid] = x;
    excl[k] = x - cnt[k];
  }
  __syncthreads();
  if (wid < UD_NCLS) {
    int sv = wtot[wid][lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, sv, o);
      if (lane >= o) sv += y;
    }
    wtot[wid][lane] = sv;                                   // inclusive over warps
    if (lane == 31) cls_total[wid] = sv;
  }
  __syncthreads();
  int base[UD_NCLS];                                        // first slot of the class: larger classes first
  int run = 0;
#pragma unroll
  for (int k = UD_NCLS - 1; k >= 0; --k) { base[k] = run; run += cls_total[k]; }
  int slot[UD_NCLS];
#pragma unroll
  for (int k = 0; k < UD_NCLS; ++k) slot[k] = base[k] + (wid > 0 ? wtot[k][wid - 1] : 0) + excl[k];
  for (int t = t0; t < t1; ++t) {
    const int2 rg = ranges[t];
    const int c = rg.y - rg.x;
    if (c <= 0) continue;
    const int v = (c + SEG - 1) / SEG;
    const int u0 = unit_start[t];
    if (u0 + v > unit_cap) continue;
    const int multi = v > 1 ? (int)0x80000000u : 0;
    for (int q = 0; q < v; ++q) {
      const int nq = min(SEG, c - q * SEG);
      const int k = (nq + UD_STEP - 1) / UD_STEP - 1;
      int pos = 0;
#pragma unroll
      for (int kk = 0; kk < UD_NCLS; ++kk)
        if (kk == k) pos = slot[kk]++;
      udesc[pos] = make_int4(t, rg.x + q * SEG, nq, (u0 + q) | multi);
    }
  }
  if (threadIdx.x == 0) counters->n_ne = run;
}

int launch_udesc(const int2* ranges, const int* unit_start, int n_tiles, int seg, int64_t unit_cap, int4* udesc,
                 Counters* counters, cudaStream_t st) {
  udesc_kernel<<<1, 1024, 0, st>>>(ranges, unit_start, n_tiles, seg, (int)unit_cap, udesc, counters);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

// ---- tile-major counting sort (order-independent blend modes) ---------------------------------
// The weighted-sum blend does not care about the order inside a tile's list, so instead of the
// multi-pass radix sort the (Gaussian,tile) pairs are grouped with ONE counting pass and ONE
// scatter pass.  `nb` persistent-style blocks each own a contiguous slice of the Gaussians and a
// shared-memory histogram over all tiles (4 B per tile):
//   cs_hist     : table[b][t] = number of pairs of block b in tile t
//   cs_colscan  : table[b][t] <- exclusive prefix over b ; total[t]
//   cs_tilescan : ranges[t], work units, counters        (one block: a few thousand tiles)
//   cs_scatter  : vals[ranges[t].x + table[b][t] + rank] = Gaussian id  (rank from a smem atomic)
// HBM traffic: 8 B/Gaussian (rect) twice + 4 B per pair written once; the 4 B x nb x tiles table
// stays in L2.  The order of ids inside (block, tile) follows the atomics, i.e. it is not
// reproducible run to run; the set is (tests compare per-tile sorted lists with the oracle).
constexpr int CS_THREADS = 1024;

__global__ void __launch_bounds__(1024)
cs_hist_kernel(int n, int per_block, int tiles_x, int n_tiles, const uint2* __restrict__ rect,
               const unsigned long long* __restrict__ tmask, const int* __restrict__ order,
               const int* __restrict__ slab_start, int* __restrict__ table) {
  extern __shared__ int hist[];
  for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) hist[t] = 0;
  __syncthreads();
  // order / slab_start: block b walks depth slab b of the Gaussians (segsort.cu); else an equal share in index order
  const int i0 = order != nullptr ? slab_start[blockIdx.x] : blockIdx.x * per_block;
  const int i1 = order != nullptr ? slab_start[blockIdx.x + 1] : min(n, i0 + per_block);
  for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    const int g = order != nullptr ? order[i] : i;
    for_each_tile(rect[g], tmask[g], tiles_x, [&](int tile) { atomicAdd(&hist[tile], 1); });
  }
  __syncthreads();
  int* dst = table + (size_t)blockIdx.x * n_tiles;
  for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) dst[t] = hist[t];
}

// table[b][t] <- exclusive prefix over b, total[t] = column sum.  A block owns 32 tiles; warp w
// scans the segment b in [w*CS_SEG, (w+1)*CS_SEG) of those columns from registers.
constexpr int CS_SEG = (CS_NB + 7) / 8;   // 37
__global__ void __launch_bounds__(256)
cs_colscan_kernel(int* __restrict__ table, int nb, int n_tiles, int* __restrict__ total) {
  __shared__ int seg[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int t = blockIdx.x * 32 + lane;
  const int b0 = w * CS_SEG;
  int v[CS_SEG];
  int sum = 0;
#pragma unroll
  for (int j = 0; j < CS_SEG; ++j) {
    const int b = b0 + j;
    v[j] = (t < n_tiles && b < nb) ? table[(size_t)b * n_tiles + t] : 0;
    sum += v[j];
  }
  seg[w][lane] = sum;
  __syncthreads();
  int run = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) run += (q < w) ? seg[q][lane] : 0;
  if (t < n_tiles) {
#pragma unroll
    for (int j = 0; j < CS_SEG; ++j) {
      const int b = b0 + j;
      if (b < nb) table[(size_t)b * n_tiles + t] = run;
      run += v[j];
    }
    if (w == 7) total[t] = run;
  }
}

// One block: exclusive scan of the per-tile totals -> ranges, counters and the work-unit table
// (same unit definition as units_kernel).  Thread t owns the contiguous chunk of tiles
// [t*per, (t+1)*per): one pass to sum its chunk, ONE block-wide scan of the 1024 chunk sums, one pass to
// write -- two barriers instead of four per 1024 tiles.
__global__ void __launch_bounds__(1024)
cs_tilescan_kernel(const int* __restrict__ total, int n_tiles, int SEG, long long max_pairs, int2* __restrict__ ranges,
                   Counters* __restrict__ counters, Counters* __restrict__ mirror, int unit_cap,
                   int* __restrict__ unit_start, int2* __restrict__ units, int4* __restrict__ udesc,
                   int* __restrict__ ne_list) {
  const int UD_STEP = (SEG + UD_NCLS - 1) / UD_NCLS;
  extern __shared__ int ts_smem[];                 // cnt[n_tiles] then ustart[n_tiles]
  int* cnt = ts_smem;
  int* ust = ts_smem + n_tiles;
  __shared__ long long wsum[32];
  __shared__ int wunits[32];
  __shared__ long long grand_s;
  __shared__ int wcls[UD_NCLS][32];               // unit descriptor table (launch_udesc): units per step class
  __shared__ int ne_count_s;
  if (threadIdx.x == 0) ne_count_s = 0;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // coalesced fetch of the per-tile totals: every load of the block is in flight at once
  for (int t = threadIdx.x; t < n_tiles; t += 1024) cnt[t] = total[t];
  __syncthreads();
  const int per = (n_tiles + 1023) / 1024;
  const int t0 = min(n_tiles, (int)threadIdx.x * per), t1 = min(n_tiles, t0 + per);
  long long c_sum = 0;
  int u_sum = 0;
  int ccnt[UD_NCLS];
#pragma unroll
  for (int kk = 0; kk < UD_NCLS; ++kk) ccnt[kk] = 0;
  for (int t = t0; t < t1; ++t) {
    const int c = cnt[t];
    c_sum += c;
    u_sum += c > 0 ? (c + SEG - 1) / SEG : 1;
    if (c > 0) {
      const int v = (c + SEG - 1) / SEG;
      ccnt[UD_NCLS - 1] += v - 1;
      const int kl = (c - (v - 1) * SEG + UD_STEP - 1) / UD_STEP - 1;
#pragma unroll
      for (int kk = 0; kk < UD_NCLS; ++kk) ccnt[kk] += (kk == kl) ? 1 : 0;
    }
  }
  int cexcl[UD_NCLS];
#pragma unroll
  for (int kk = 0; kk < UD_NCLS; ++kk) {
    int z = ccnt[kk];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int zz = __shfl_up_sync(0xffffffffu, z, o);
      if (lane >= o) z += zz;
    }
    if (lane == 31) wcls[kk][wid] = z;
    cexcl[kk] = z - ccnt[kk];
  }
  // inclusive warp scans of (pairs, units)
  long long x = c_sum;
  int y = u_sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long xx = __shfl_up_sync(0xffffffffu, x, o);
    const int yy = __shfl_up_sync(0xffffffffu, y, o);
    if (lane >= o) { x += xx; y += yy; }
  }
  if (lane == 31) { wsum[wid] = x; wunits[wid] = y; }
  __syncthreads();
  if (wid == 0) {
    long long a = wsum[lane];
    int b = wunits[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long aa = __shfl_up_sync(0xffffffffu, a, o);
      const int bb = __shfl_up_sync(0xffffffffu, b, o);
      if (lane >= o) { a += aa; b += bb; }
    }
    wsum[lane] = a;
    wunits[lane] = b;
    if (lane == 31) grand_s = a;
  } else if (wid <= UD_NCLS) {
    int z = wcls[wid - 1][lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int zz = __shfl_up_sync(0xffffffffu, z, o);
      if (lane >= o) z += zz;
    }
    wcls[wid - 1][lane] = z;                       // inclusive over warps
  }
  __syncthreads();
  const long long grand = grand_s;
  const bool ov = grand > max_pairs;              // on overflow nothing is rendered: kept = 0, every range empty
  long long start = (wid > 0 ? wsum[wid - 1] : 0) + (x - c_sum);
  int ustart = (wid > 0 ? wunits[wid - 1] : 0) + (y - u_sum);
  if (ov) ustart = t0;                            // every tile keeps exactly one (empty) unit
  // chunk pass in shared memory: cnt[t] <- start of the tile's range (its count is recovered from the next
  // start), ust[t] <- first unit
  int nunits = ov ? n_tiles : wunits[31];
  // descriptor slots of this thread's units: classes laid out largest first, tile order inside a class
  int slot[UD_NCLS];
  int n_ne = 0;
#pragma unroll
  for (int kk = UD_NCLS - 1; kk >= 0; --kk) {
    slot[kk] = n_ne + (wid > 0 ? wcls[kk][wid - 1] : 0) + cexcl[kk];
    n_ne += wcls[kk][31];
  }
  for (int t = t0; t < t1; ++t) {
    const int c = ov ? 0 : cnt[t];
    ust[t] = ustart;
    const int v = c > 0 ? (c + SEG - 1) / SEG : 1;
    if (c > 0 && ustart + v <= unit_cap) {
      const int multi = v > 1 ? (int)0x80000000u : 0;
      for (int q = 0; q < v; ++q) {
        const int nq = min(SEG, c - q * SEG);
        const int kq = (nq + UD_STEP - 1) / UD_STEP - 1;
        int pos = 0;
#pragma unroll
        for (int kk = 0; kk < UD_NCLS; ++kk)
          if (kk == kq) pos = slot[kk]++;
        udesc[pos] = make_int4(t, (int)start + q * SEG, nq, (ustart + q) | multi);
      }
    }
    ustart += v;
    cnt[t] = (int)start;
    start += c;
  }
  __syncthreads();
  // coalesced write-out
  const int kept = ov ? 0 : (int)grand;
  for (int t = threadIdx.x; t < n_tiles; t += 1024) {
    const int s0 = cnt[t], s1 = (t + 1 < n_tiles) ? cnt[t + 1] : kept;
    const int c = ov ? 0 : s1 - s0;
    ranges[t] = c > 0 ? make_int2(s0, s1) : make_int2(0, 0);   // empty tiles read (0,0) like the radix path
    if (ne_list != nullptr && c > 1) ne_list[1 + atomicAdd(&ne_count_s, 1)] = t;   // tiles with something to order
    const int u0 = ust[t];
    unit_start[t] = u0;
    const int v = c > 0 ? (c + SEG - 1) / SEG : 1;
    for (int q = 0; q < v; ++q)
      if (u0 + q < unit_cap) units[u0 + q] = make_int2(t, q);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (ne_list != nullptr) ne_list[0] = ne_count_s;
    counters->needed = grand;
    counters->kept = kept;
    counters->overflow = ov ? 1 : 0;
    counters->n_ne = ov ? 0 : n_ne;
    if (mirror != nullptr) {
      mirror->needed = grand;
      mirror->kept = kept;
      mirror->overflow = ov ? 1 : 0;
    }
    unit_start[n_tiles] = nunits < unit_cap ? nunits : unit_cap;
  }
}

__global__ void __launch_bounds__(1024)
cs_scatter_kernel(int n, int per_block, int tiles_x, int n_tiles, const uint2* __restrict__ rect,
                  const unsigned long long* __restrict__ tmask, const int* __restrict__ order,
                  const int* __restrict__ slab_start, const int* __restrict__ table,
                  const int2* __restrict__ ranges, const Counters* __restrict__ counters, int* __restrict__ vals) {
  extern __shared__ int off[];
  if (counters->overflow) return;
  const int* src = table + (size_t)blockIdx.x * n_tiles;
  for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) off[t] = ranges[t].x + src[t];
  __syncthreads();
  const int i0 = order != nullptr ? slab_start[blockIdx.x] : blockIdx.x * per_block;
  const int i1 = order != nullptr ? slab_start[blockIdx.x + 1] : min(n, i0 + per_block);
  for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    const int g = order != nullptr ? order[i] : i;
    for_each_tile(rect[g], tmask[g], tiles_x, [&](int tile) {
      const int pos = atomicAdd(&off[tile], 1);
      vals[pos] = g;
    });
  }
}

bool counting_sort_fits(int n_tiles) { return (size_t)n_tiles * 8 <= CS_MAX_SMEM; }   // tile scan stages 2 ints per tile
int counting_sort_blocks(int n) {
  int nb = (n + CS_THREADS - 1) / CS_THREADS;
  const int cap = 2 * sm_count() < CS_NB ? 2 * sm_count() : CS_NB;   // two blocks per SM
  return nb < 1 ? 1 : (nb > cap ? cap : nb);
}

int launch_counting_sort(const ViewParams& vp, int n, int64_t max_pairs, const uint2* rect,
                         const unsigned long long* tmask, const int* order, const int* slab_start, int* table, int* total,
                         int2* ranges, Counters* counters, Counters* mirror, int64_t unit_cap, int* unit_start, int2* units,
                         int4* udesc, int* ne_list, int* vals, int stage, cudaStream_t st) {
  const size_t smem = (size_t)vp.n_tiles * 4;
  B2S_CUDA_TRY(per_device_once(ONCE_COUNTING_SORT, [] {
    cudaError_t e = cudaFuncSetAttribute(cs_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CS_MAX_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(cs_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CS_MAX_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(cs_tilescan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CS_MAX_SMEM);
    return e;
  }));
  const int nb = counting_sort_blocks(n);
  const int per_block = (n + nb - 1) / nb;
  static const int threads = [] { const char* e = getenv("B2S_CS_THREADS"); const int v = e ? atoi(e) : 0;
                                  return (v == 256 || v == 512 || v == 1024) ? v : CS_THREADS; }();
  if (stage == 0) {
    cs_hist_kernel<<<nb, threads, smem, st>>>(n, per_block, vp.tiles_x, vp.n_tiles, rect, tmask, order, slab_start, table);
    B2S_LAUNCH_CHECK();
    cs_colscan_kernel<<<(vp.n_tiles + 31) / 32, 256, 0, st>>>(table, nb, vp.n_tiles, total);
    B2S_LAUNCH_CHECK();
    cs_tilescan_kernel<<<1, 1024, 2 * smem, st>>>(total, vp.n_tiles, vp.seg, (long long)max_pairs, ranges, counters, mirror,
                                           (int)unit_cap, unit_start, units, udesc, ne_list);
    B2S_LAUNCH_CHECK();
  } else {
    cs_scatter_kernel<<<nb, threads, smem, st>>>(n, per_block, vp.tiles_x, vp.n_tiles, rect, tmask, order, slab_start, table, ranges, counters,
                                                    vals);
    B2S_LAUNCH_CHECK();
  }
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
