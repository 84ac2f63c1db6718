// Shared device/host definitions for libb2splat (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>

#include "b2splat.h"

namespace b2s {

constexpr int TILE = B2S_TILE;        // 16x16 pixel tiles
constexpr int TILE_PIX = TILE * TILE;
constexpr int REC_F4 = 3;             // per-Gaussian blend record: 3 x float4 = 48 B
constexpr int GACC_F = 12;            // per-Gaussian backward accumulator: 12 floats = 48 B
constexpr int GBUF_FRAG_WORDS = 80 * 32;   // per-tile g-buffer as fp16 hi/lo MMA fragments: 80 registers x 32 lanes

// Per-view constants handed to every kernel by value.
struct ViewParams {
  float view[16];
  float proj[16];
  float cam[3];     // inv(view)[:3,3]  (torch_renderer.py:81-83)
  float bg[3];
  float k;          // cutoff in sigmas
  float wm1, hm1;   // float(W-1), float(H-1)
  float wf, hf;     // float(W), float(H)
  float fx, fy;     // |P00|, |P11|
  int width, height, tiles_x, tiles_y, n_tiles;
  int style, sh, act, exact_bbox, mode;
  int seg;          // Gaussians per work unit of the blend kernels (unit_size(): grows with the image)
  int keep_depth;   // accumulate the depth plane D in the saved accumulators even without a depth image (fit loop with a depth loss)
  const float* bg_dev;   // optional DEVICE pointer to 3 floats overriding bg[] (b2s_params.background_dev): the
                         // drop-in receives the background as a device tensor and must not sync to read it
};

// background colour of the view: the device override when given, else the host-supplied constants
__device__ __forceinline__ float view_bg(const ViewParams& vp, int c) {
  return vp.bg_dev != nullptr ? __ldg(vp.bg_dev + c) : vp.bg[c];
}

// One Gaussian after projection.  Every quantity that feeds an integer (bbox, tile rect,
// depth key) is computed with explicitly rounded, non-contracted fp32 ops in exactly the
// order of oracle/bins_oracle.c so that GPU and CPU agree bit for bit.
struct Proj {
  float px, py, sx, sy, zabs, zcam;
  float w, wsafe, ndcx, ndcy;   // for the backward chain
  float ax, ay;                 // unclamped sigmas
  int xmin, ymin, xmax, ymax;   // pixel bbox (inclusive)
  bool ok;
  bool valid;                   // ok before the bbox-on-screen test (extension modes with their own footprint: splat2d.cu)
};

__device__ __forceinline__ float dot4_lr(const float* m, float a, float b, float c, float d) {
  float t = __fmul_rn(m[0], a);
  t = __fadd_rn(t, __fmul_rn(m[1], b));
  t = __fadd_rn(t, __fmul_rn(m[2], c));
  t = __fadd_rn(t, __fmul_rn(m[3], d));
  return t;
}

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
// torch.nn.functional.softplus (beta=1, threshold=20)
__device__ __forceinline__ float softplusf_acc(float x) { return x > 20.0f ? x : log1pf(expf(x)); }

__device__ __forceinline__ Proj project_gaussian(const ViewParams& vp, float mx, float my, float mz,
                                                 float s0, float s1, float op) {
  Proj r;
  float cam[4], clip[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) cam[q] = dot4_lr(vp.view + 4 * q, mx, my, mz, 1.0f);
#pragma unroll
  for (int q = 0; q < 4; ++q) clip[q] = dot4_lr(vp.proj + 4 * q, cam[0], cam[1], cam[2], cam[3]);
  const float w = clip[3];
  float nx, ny, nz, ssx, ssy;
  bool ok;
  if (vp.style == B2S_STYLE_TORCH) {
    const float ws = (fabsf(w) < 1e-8f) ? 1.0f : w;
    nx = __fdiv_rn(clip[0], ws);
    ny = __fdiv_rn(clip[1], ws);
    nz = __fdiv_rn(clip[2], ws);
    ok = (nz >= -1.0f) && (nz <= 1.0f) && (w != 0.0f);
    r.zabs = fmaxf(fabsf(cam[2]), 1e-6f);
    ssx = fabsf(s0);
    ssy = fabsf(s1);
    ok = ok && (op >= 0.0f);   // op == 0 stays: weight 0, but clamp_min(0) passes dL/dop at 0
    r.wsafe = ws;
  } else {
    const float iw = __fdiv_rn(1.0f, (w == 0.0f) ? 1.0f : w);
    nx = __fmul_rn(clip[0], iw);
    ny = __fmul_rn(clip[1], iw);
    nz = __fmul_rn(clip[2], iw);
    ok = (w != 0.0f) && !(nz < -1.0f || nz > 1.0f) && (nz == nz);
    r.zabs = __fadd_rn(fabsf(cam[2]), 1e-6f);
    ssx = s0;
    ssy = s1;
    ok = ok && (op >= 1e-5f);
    r.wsafe = (w == 0.0f) ? 1.0f : w;
  }
  r.valid = ok;
  r.w = w;
  r.ndcx = nx;
  r.ndcy = ny;
  r.zcam = cam[2];
  r.px = __fmul_rn(__fadd_rn(__fmul_rn(nx, 0.5f), 0.5f), vp.wm1);
  r.py = __fmul_rn(__fsub_rn(1.0f, __fadd_rn(__fmul_rn(ny, 0.5f), 0.5f)), vp.hm1);
  r.ax = __fdiv_rn(__fmul_rn(__fmul_rn(__fmul_rn(ssx, 0.5f), vp.wf), vp.fx), r.zabs);
  r.ay = __fdiv_rn(__fmul_rn(__fmul_rn(__fmul_rn(ssy, 0.5f), vp.hf), vp.fy), r.zabs);
  r.sx = fmaxf(r.ax, 1.0f);
  r.sy = fmaxf(r.ay, 1.0f);
  const float rx = __fmul_rn(vp.k, r.sx), ry = __fmul_rn(vp.k, r.sy);
  const float lox = floorf(__fsub_rn(r.px, rx)), hix = ceilf(__fadd_rn(r.px, rx));
  const float loy = floorf(__fsub_rn(r.py, ry)), hiy = ceilf(__fadd_rn(r.py, ry));
  ok = ok && (hix >= 0.0f) && (lox <= vp.wm1) && (hiy >= 0.0f) && (loy <= vp.hm1);
  r.ok = ok;
  if (ok) {
    r.xmin = (int)fmaxf(lox, 0.0f);
    r.xmax = (int)fminf(hix, vp.wm1);
    r.ymin = (int)fmaxf(loy, 0.0f);
    r.ymax = (int)fminf(hiy, vp.hm1);
  } else {
    r.xmin = r.ymin = 0;
    r.xmax = r.ymax = -1;
  }
  return r;
}

// ascending key order == descending camera z; same as b2o_depth_bits (oracle/bins_oracle.c)
__device__ __forceinline__ uint32_t depth_bits(float zcam) {
  uint32_t u = __float_as_uint(zcam);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ~u;
}

// Tile culling for the torch-style weighted sum.  A tile of the k-sigma bbox rect is kept iff the pixel centre of
// the tile nearest to the Gaussian lies inside the k-sigma ELLIPSE (every pixel of a dropped tile has weight
// < op * exp(-k^2/2), the same bound the bbox itself enforces along the axes).  Bit (ty-ty0)*w + (tx-tx0) of the
// result, for rects of at most 8 x 8 tiles; explicitly rounded ops in the order of oracle/bins_oracle.c
// (b2o_tile_mask) so that GPU and CPU agree bit for bit.
__device__ __forceinline__ unsigned long long tile_cull_mask(float px, float py, float sx, float sy, float k, int tx0,
                                                             int ty0, int w, int h, bool cull) {
  const float isx = __fdiv_rn(1.0f, sx), isy = __fdiv_rn(1.0f, sy), kk = __fmul_rn(k, k);
  float uxx[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float lo = (float)((tx0 + c) * TILE) + 0.5f, hi = lo + (float)(TILE - 1);
    const float u = __fmul_rn(__fsub_rn(fminf(fmaxf(px, lo), hi), px), isx);
    uxx[c] = __fmul_rn(u, u);
  }
  // columns beyond the rect never pass: +inf fails the test below, and without culling the row is just the rect width
  const uint32_t wmask = (1u << w) - 1u;
#pragma unroll
  for (int c = 0; c < 8; ++c) uxx[c] = (c < w) ? uxx[c] : INFINITY;
  unsigned long long m = 0ull;
  for (int r = 0; r < h; ++r) {
    const float lo = (float)((ty0 + r) * TILE) + 0.5f, hi = lo + (float)(TILE - 1);
    const float u = __fmul_rn(__fsub_rn(fminf(fmaxf(py, lo), hi), py), isy);
    const float uyy = __fmul_rn(u, u);
    uint32_t row = 0;                               // the row's columns as 8 bits, placed with ONE 64-bit shift
#pragma unroll
    for (int c = 0; c < 8; ++c) row |= (__fadd_rn(uxx[c], uyy) <= kk) ? (1u << c) : 0u;
    row = cull ? row : wmask;
    m |= (unsigned long long)row << (r * w);
  }
  return m;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// The 16 factors of one axis of one Gaussian over a tile, f[c] = 2^(q (d0 + c)^2 + e0), c = 0..15, returned as pairs
// out[j] = (f[2j], f[2j+1]).  Evaluated by recurrence from the middle pair (columns 8, 9) outwards in steps of two
// columns: f[c+2]/f[c] = 2^(q (4 d + 4)), f[c-2]/f[c] = 2^(q (4 - 4 d)), and consecutive ratios differ by the constant
// 2^(8 q) -- 7 MUFU.EX2 + 13 packed multiplies instead of 16 MUFU.EX2 + 16 FMUL + 16 FFMA.  Every intermediate is a
// true factor value or a ratio of two of them (|exponent| <= 4 * 0.72 * (cutoff + 8) < 128 because sigma >= 1 px), so
// nothing overflows; relative error <= (1 + 4 + 6) ulp(ex2.approx) ~ 2.6e-6.  A middle value that underflows (the
// Gaussian is > 13 sigma from the tile centre) zeroes the whole axis: the largest value lost is exp(-24) at cutoff 7.
__device__ __forceinline__ void factors16(float q, float d0, float e0, float2 (&out)[8]) {
  const float d8 = d0 + 8.0f, d9 = d0 + 9.0f;
  const float2 d = make_float2(d8, d9);
  const float2 e = __ffma2_rn(__fmul2_rn(make_float2(q, q), d), d, make_float2(e0, e0));
  const float q4 = 4.0f * q;
  float2 Ru = make_float2(ex2_approx(q4 * (d8 + 1.0f)), ex2_approx(q4 * (d9 + 1.0f)));
  float2 Rd = make_float2(ex2_approx(q4 * (1.0f - d8)), ex2_approx(q4 * (1.0f - d9)));
  const float k = ex2_approx(8.0f * q);
  const float2 K = make_float2(k, k);
  float2 U = make_float2(ex2_approx(e.x), ex2_approx(e.y)), D = U;
  out[4] = U;
#pragma unroll
  for (int j = 5; j < 8; ++j) {
    U = __fmul2_rn(U, Ru);
    out[j] = U;
    if (j < 7) Ru = __fmul2_rn(Ru, K);
  }
#pragma unroll
  for (int j = 3; j >= 0; --j) {
    D = __fmul2_rn(D, Rd);
    out[j] = D;
    if (j > 0) Rd = __fmul2_rn(Rd, K);
  }
}

// ---- state / workspace layout (all offsets 256-B aligned) -------------------------------
struct StateLayout {
  size_t rec, ranges, vals, acc, counters, unit_start, units, udesc, cmask, total;
};
struct WorkLayout {
  size_t rect, tmask, dbits, cnt, bsum, keysA, keysB, valsB, hist, hsum, gacc, partial, gbuf, cs_table, cs_total;
  size_t oorder, oslab, osmall;   // depth slabs of the Gaussians (segsort.cu): order[n], slab id[n], splitters / counters
  size_t ne_list, rest_list;      // compact work lists: non-empty tiles (group sort), live segments (sorted blend)
  size_t total;
};
// A work unit of the blend kernels: one tile x one segment of at most SEG Gaussians of its
// list.  Splitting long lists keeps the units uniform (a 1080p tile of the C4 scene holds up
// to ~5000 Gaussians; one CTA per tile left the SMs idle a third of the time).
// Unit size: long units amortise the per-unit costs of the persistent tcgen05 kernels (accumulator read-back,
// partial planes, plane fetch) -- 512 -> 4096 takes the C4 forward from 19.8 to 17.1 ms -- but a unit is the grain of
// the static load balance, so an image needs enough tiles to keep every CTA busy with several of them.
constexpr int SEG_MIN = 512;
inline int unit_size(int n_tiles) { return n_tiles >= 4096 ? 4096 : (n_tiles >= 1024 ? 1024 : SEG_MIN); }
inline int64_t max_units(int width, int height, int64_t max_pairs) {
  const int64_t tiles = (int64_t)((width + TILE - 1) / TILE) * ((height + TILE - 1) / TILE);
  return (max_pairs > 0 ? max_pairs : 0) / SEG_MIN + tiles;   // sum_t max(1, ceil(c_t/seg)) <= P1/seg + tiles, seg >= SEG_MIN
}
constexpr int SORT_KPB = 4096;   // keys per radix block
constexpr int CS_NB = 296;       // counting-sort blocks: 2 per SM, at most this many (the table is sized for it)
constexpr size_t CS_MAX_SMEM = 200 * 1024;   // per-block tile histogram (4 B per tile) must fit
constexpr int PRE_BLOCK = 256;   // Gaussians per preprocess block

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

inline StateLayout state_layout(int n, int width, int height, int64_t max_pairs) {
  StateLayout L;
  const int tiles = ((width + TILE - 1) / TILE) * ((height + TILE - 1) / TILE);
  size_t o = 0;
  L.counters = o; o += align_up(64);
  L.rec = o;      o += align_up((size_t)(n > 0 ? n : 1) * REC_F4 * 16);
  L.ranges = o;   o += align_up((size_t)tiles * 8);
  L.vals = o;     o += align_up((size_t)(max_pairs > 0 ? max_pairs : 1) * 4);
  L.acc = o;      o += align_up((size_t)width * height * 5 * 4);
  L.unit_start = o; o += align_up((size_t)(tiles + 1) * 4);
  L.units = o;    o += align_up((size_t)max_units(width, height, max_pairs) * 8);
  L.udesc = o;    o += align_up((size_t)max_units(width, height, max_pairs) * 16);   // non-empty units, largest first
  L.cmask = o;    o += align_up((size_t)(n > 0 ? n : 1));   // colour clamp mask, 1 B per Gaussian (torch_renderer.py:144)
  L.total = o;
  return L;
}

inline WorkLayout work_layout(int n, int width, int height, int64_t max_pairs) {
  WorkLayout L;
  const size_t nn = (size_t)(n > 0 ? n : 1);
  const size_t mp = (size_t)(max_pairs > 0 ? max_pairs : 1);
  const size_t nb_pre = (nn + PRE_BLOCK - 1) / PRE_BLOCK;
  const size_t nb_sort = (mp + SORT_KPB - 1) / SORT_KPB;
  size_t o = 0;
  L.rect = o;  o += align_up(nn * 8);
  L.tmask = o; o += align_up(nn * 8);   // kept tiles of rects <= 8x8 (tile_cull_mask)
  L.dbits = o; o += align_up(nn * 4);
  L.cnt = o;   o += align_up(nn * 4);
  L.bsum = o;  o += align_up((nb_pre + 1) * 8);
  L.keysA = o; o += align_up(mp * 8);
  L.keysB = o; o += align_up(mp * 8);
  L.valsB = o; o += align_up(mp * 4);
  L.hist = o;  o += align_up(nb_sort * 256 * 4);
  L.hsum = o;  o += align_up(((nb_sort * 256 + 4095) / 4096 + 1) * 4);
  L.gacc = o;  o += align_up(nn * GACC_F * 4);
  const size_t tiles = (size_t)((width + TILE - 1) / TILE) * ((height + TILE - 1) / TILE);
  L.partial = o; o += align_up((size_t)max_units(width, height, max_pairs) * 5 * TILE_PIX * 4);
  L.gbuf = o;  o += align_up(tiles * GBUF_FRAG_WORDS * 4) + align_up(tiles * 4 + 64);   // plane fragments + per-tile scale + depth statistics
  L.cs_table = o; o += align_up((size_t)CS_NB * tiles * 4);
  L.cs_total = o; o += align_up(tiles * 4);
  L.oorder = o; o += align_up(nn * 4);
  L.oslab = o;  o += align_up(nn * 4);
  L.osmall = o; o += align_up(4 * 1024 * 4);   // splitters | counts | slab_start | cursor, 1024 words each
  L.ne_list = o; o += align_up((tiles + 1) * 4);
  L.rest_list = o; o += align_up(((size_t)max_units(width, height, max_pairs) + 1) * 8);
  L.total = o;
  return L;
}

// Per-view block written by the batched preprocess (b2s_preprocess_views): what launch_preprocess leaves in
// state (rec, cmask) and workspace (rect, tmask) for the counting-sort path, for one view.
struct PreparedLayout {
  size_t rec, cmask, rect, tmask, total;
};
inline PreparedLayout prepared_layout(int n) {
  PreparedLayout L;
  const size_t nn = (size_t)(n > 0 ? n : 1);
  size_t o = 0;
  L.rec = o;   o += align_up(nn * REC_F4 * 16);
  L.cmask = o; o += align_up(nn);
  L.rect = o;  o += align_up(nn * 8);
  L.tmask = o; o += align_up(nn * 8);
  L.total = o;
  return L;
}

// counters block at the head of the state buffer
struct Counters {
  long long needed;   // pairs the view produces
  int kept;           // pairs actually emitted (<= max_pairs)
  int overflow;       // 1 if needed > max_pairs
  int n_ne;           // non-empty work units (entries of the unit descriptor table)
  int pad_[3];
};

// ---- host-side error plumbing ---------------------------------------------------------------
void set_error(const char* fmt, ...);
#define B2S_CUDA_TRY(expr)                                                               \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      b2s::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return B2S_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)
void count_launch();
// launches of the weighted-sum blend kernel families (b2s_path_counts)
enum Path { PATH_FWD_UMMA = 0, PATH_FWD_OTHER, PATH_BWD_UMMA, PATH_BWD_OTHER, PATH_COUNT };
void count_path(int which);
// SM count of the current device (cudaDevAttrMultiProcessorCount, cached per device): persistent grids and the
// counting-sort block count are sized from it, nothing assumes 148
int sm_count();
// One-time per-device setup (cudaFuncSetAttribute ...): function attributes are per device and a process may drive
// several GPUs from several host threads, so the guard is a std::once_flag per (slot, current device).
enum OnceSlot { ONCE_FWD_UMMA = 0, ONCE_COUNTING_SORT, ONCE_BWD_UMMA, ONCE_SORT, ONCE_SORTED_BLEND, ONCE_SLOTS };
std::once_flag& device_once_flag(int slot);
template <class F>
inline cudaError_t per_device_once(int slot, F f) {
  cudaError_t e = cudaSuccess;
  std::call_once(device_once_flag(slot), [&] { e = f(); });
  return e;
}
#define B2S_LAUNCH_CHECK()             \
  do {                                 \
    b2s::count_launch();               \
    B2S_CUDA_TRY(cudaGetLastError());  \
  } while (0)

// pipeline stages, for the optional CUDA-event timing of b2s_timing_*
enum Stage { ST_PREPROCESS = 0, ST_BIN, ST_SORT, ST_RANGES, ST_BLEND_FWD, ST_LOSS, ST_BLEND_BWD, ST_PREPROCESS_BWD, ST_ADAM, ST_COUNT };

// ---- kernel launchers (defined in the .cu files) ------------------------------------------
int launch_preprocess(const ViewParams& vp, const float* means, const float* scales, const float* colors,
                      const float* opac, int n, float4* rec, uint8_t* cmask, uint2* rect, unsigned long long* tmask,
                      uint32_t* dbits, int* cnt, long long* bsum, float* dbg /*px,py,sx,sy,zabs planes or null*/,
                      int* dbg_bbox, cudaStream_t st);
int launch_preprocess_views(const ViewParams* views_dev, int num_views, int sh, const float* means, const float* scales,
                            const float* colors, const float* opac, int n, char* prepared, cudaStream_t st);
// mirror: optional second destination of the pair counters (pinned host memory, see b2s_ticket_info)
int launch_bin(const ViewParams& vp, int n, int64_t max_pairs, const uint2* rect, const unsigned long long* tmask,
               const uint32_t* dbits,
               const int* cnt, long long* bsum, unsigned long long* keys, int* vals, Counters* counters,
               Counters* mirror, cudaStream_t st);
// sorts (keysA, valsA) on key bits [begin_bit,end_bit); returns via *result_in_B where the result lives
int launch_sort(unsigned long long* keysA, int* valsA, unsigned long long* keysB, int* valsB, int64_t cap,
                const int* count_dev, int begin_bit, int end_bit, int* hist, int* hsum, int* result_in_B,
                cudaStream_t st);
inline int sort_passes(int begin_bit, int end_bit) { return end_bit > begin_bit ? (end_bit - begin_bit + 7) / 8 : 0; }
int launch_ranges(const unsigned long long* keys, const int* count_dev, int64_t cap, int n_tiles, int2* ranges,
                  cudaStream_t st);
// tile-major counting sort (bin.cu): stage 0 = histogram + scans (ranges, units, counters), stage 1 = scatter
bool counting_sort_fits(int n_tiles);
int counting_sort_blocks(int n);
// order / slab_start: optional depth slabs of the Gaussians (segsort.cu): block b walks order[slab_start[b] .. slab_start[b+1])
int launch_counting_sort(const ViewParams& vp, int n, int64_t max_pairs, const uint2* rect,
                         const unsigned long long* tmask, const int* order, const int* slab_start, int* table, int* total,
                         int2* ranges, Counters* counters, Counters* mirror, int64_t unit_cap, int* unit_start, int2* units,
                         int4* udesc, int* ne_list /* [0] = count, then the non-empty tiles; may be null */, int* vals,
                         int stage, cudaStream_t st);
int launch_units(const int2* ranges, int n_tiles, int seg, int64_t unit_cap, int* unit_start, int2* units, cudaStream_t st);
// Unit descriptor table of the persistent tcgen05 blend kernels: the NON-EMPTY units as {tile, first pair, pairs,
// unit index | multi-unit-tile flag << 31}, ordered by their number of 128-Gaussian steps, largest first, so that CTA i
// taking entries i, i + grid, i + 2 grid, ... gets the same mix of work as every other CTA (static, balanced), and a
// unit costs ONE 16-byte load instead of the dependent units -> ranges pair.
int launch_udesc(const int2* ranges, const int* unit_start, int n_tiles, int seg, int64_t unit_cap, int4* udesc,
                 Counters* counters, cudaStream_t st);
int launch_blend_wsum_fwd(const ViewParams& vp, const float4* rec, const int* vals, const int2* ranges,
                          const int* unit_start, const int2* units, const int4* udesc, const Counters* counters,
                          int64_t unit_cap, float* partial,
                          float* out_rgb, float* out_alpha, float* out_depth, float* acc, uint8_t* out_rgba,
                          cudaStream_t st);
// sat_seg: n_tiles ints of scratch (first segment of a tile that saturates every pixel on its own)
int launch_blend_sorted_fwd(const ViewParams& vp, const float4* rec, const int* vals, const int2* ranges,
                            const int* unit_start, float* partial, int* sat_seg, int2* rest_list, float* out_rgb,
                            float* out_alpha, uint8_t* out_rgba, Counters* dbg /* -DB2S_STATS builds only */, cudaStream_t st);
// depth order inside the tiles (segsort.cu): the Gaussians partitioned into nb depth slabs (order[], slab_start[]);
// after the slab-wise counting sort, launch_group_sort orders the inside of every (block, tile) group.
// scratch: max_pairs x 8 B; keys_out: optional rebuilt sorted keys (tile << 32 | depth bits) for b2s_dump_bins
int launch_depth_slabs(const uint32_t* dbits, int n, int nb, uint32_t* splitters, int* count, int* slab_start, int* cursor,
                       int* slab_id, int* order, cudaStream_t st);
int launch_group_sort(const ViewParams& vp, const int* table, const int* total, const int2* ranges, const int* ne_list, int nb,
                      const uint32_t* dbits, const Counters* counters, int* vals, unsigned long long* scratch,
                      unsigned long long* keys_out, cudaStream_t st);
int launch_gacc_init(const uint8_t* cmask, float* gacc, int n, cudaStream_t st);
// fl != null: the image gradients are those of the fit loss, evaluated from the saved accumulators inside the
// g-buffer kernel (g_rgb / g_alpha / g_depth are ignored)
struct FitLossArgs {
  const float* tgt;
  const float* mask;     // may be null
  float w_sil, scale;
  float* loss_accum;
  const float* depth_gt; // may be null: the depth term  w_depth * mean|depth/(max depth + 1e-6) - depth_gt|
  float w_depth;
  // 8-bit target / mask (the decoded image bytes, value / 255 as np.asarray(img, float32) / 255 of
  // fit_multiview_stub.py:16-23): when non-null they replace tgt / mask and the loss kernel converts on the fly
  const uint8_t* tgt_u8;
  const uint8_t* mask_u8;
};
int launch_blend_wsum_bwd(const ViewParams& vp, const float4* rec, const int* vals, const int2* ranges,
                          const int* unit_start, const int2* units, const int4* udesc, const Counters* counters,
                          int64_t unit_cap, const float* acc,
                          const float* g_rgb, const float* g_alpha, const float* g_depth, const FitLossArgs* fl,
                          float* gbuf, float* gacc, cudaStream_t st);
// single != null: one view passed by value (views_dev must be null, num_views 1, gacc = n x 12 floats);
// else views_dev[num_views] in device memory and gacc = num_views x n x 12 floats.
int launch_preprocess_bwd(const ViewParams* single, const ViewParams* views_dev, int num_views, int sh,
                          const float* means, const float* scales, const float* colors, const float* opac, int n,
                          int first, int count /* Gaussians [first, first + count) */, const float* gacc, float* g_means,
                          float* g_scales, float* g_colors, float* g_opac, int accumulate, cudaStream_t st);
// extension modes (splat2d.cu): rotations + EWA covariance, differentiable front-to-back compositing
int launch_ext_preprocess(const ViewParams& vp, const float* means, const float* scales, const float* rotations /* (N,4) or null */,
                          const float* colors, const float* opac, int n, float dilation, float4* rec, uint8_t* cmask,
                          uint2* rect, unsigned long long* tmask, uint32_t* dbits, int* cnt, long long* bsum,
                          cudaStream_t st);
int launch_blend_ext_fwd(const ViewParams& vp, int over, const float4* rec, const int* vals, const int2* ranges,
                         float* out_rgb, float* out_alpha, float* out_depth, float* acc, cudaStream_t st);
int launch_blend_ext_bwd(const ViewParams& vp, int over, const float4* rec, const int* vals, const int2* ranges,
                         const float* acc, const float* g_rgb, const float* g_alpha, const float* g_depth, float* gacc,
                         int n, cudaStream_t st);
int launch_ext_bwd(const ViewParams& vp, const float* means, const float* scales, const float* rotations,
                   const float* colors, const float* opac, int n, float dilation, const float* gacc, const uint8_t* cmask,
                   float* g_means, float* g_scales, float* g_rot, float* g_colors, float* g_opac, cudaStream_t st);
int launch_scan_i32(int* data, int len, int* bs, cudaStream_t st);
size_t densify_workspace_bytes(int n);
int launch_densify_prune(const float* means, const float* scales_raw, const float* op_raw, const float* colors, int n,
                         int color_floats, int max_gaussians, double ratio, float prune_opacity, unsigned long long seed,
                         unsigned long long iter, float* o_means, float* o_scales, float* o_op, float* o_colors,
                         int* n_new_dev, void* ws, cudaStream_t st);
int launch_fit_loss(const float* rgb, const float* alpha, const float* tgt, const float* mask, int width,
                    int height, float w_sil, float scale, float* g_rgb, float* g_alpha, float* loss_accum,
                    cudaStream_t st);
int launch_u8_to_f32(const uint8_t* src, float* dst, int64_t count, cudaStream_t st);
int launch_adam(float* params, const float* grads, float* m, float* v, int64_t count, int step, float lr,
                float b1, float b2, float eps, int64_t sb, int64_t se, float reg_scale, int64_t ob, int64_t oe,
                float reg_op, const float* skip_flag, int* skipped_count, cudaStream_t st);

int launch_adam_multimem(float* params_mc, const float* grads_mc, const float* params_local, float* m, float* v,
                         int64_t count, int rank, int world, int step, float lr, float b1, float b2, float eps, int64_t sb,
                         int64_t se, float reg_scale, int64_t ob, int64_t oe, float reg_op, const float* skip_flag,
                         int* skipped_count, cudaStream_t st);
int launch_tail_multimem(const float* tail_mc, float* out, int count, cudaStream_t st);
void multimem_share(int64_t count, int rank, int world, long long* lo, long long* hi);

}  // namespace b2s
