// Extension modes named by `north_star` that the reference does not have:
//   * `rotations` (N,4) + EWA 2-D covariance: the Gaussian is a full 3-D covariance R diag(s^2) R^T projected through
//     the exact Jacobian of the reference's pixel projection (python/torch_renderer.py:57-78); its screen footprint
//     has a dx*dy cross term, so the separable tensor-core contraction of blend_fwd.cu / blend_bwd.cu does not apply
//     and the blend runs per pixel on the FP32 pipe;
//   * a DIFFERENTIABLE front-to-back "over" compositing with the rule of the reference CPU renderer's sorted mode
//     (src/renderer_cpu.cpp:196-215: exact k-sigma pixel bbox, a < 1e-5 skipped, contrib = T a).
// Both are pinned by oracle/ext_oracle.py (dense torch, autograd in float64), not by the reference.
//
// One record format serves both projections (reference "billboard" sigmas or EWA conic), both blends read it:
//   rec[3i+0] = {px, py, qxx, qyy}          w = op * 2^(qxx dx^2 + qyy dy^2 + qxy dx dy),
//   rec[3i+1] = {qxy, op, zabs, bbox x}     qxx = -log2(e)/2 * A, qyy = -log2(e)/2 * C, qxy = -log2(e) * B (conic A,B,C)
//   rec[3i+2] = {r, g, b, bbox y}           bbox = min | max << 16 (pixels, inclusive)
// The front end (tile rects, 64-bit tile|depth order, tile ranges) is the shared one of bin.cu / sort.cu / segsort.cu.
// Backward: per-pixel reverse compositing per tile, gradients reduced over the warp with shuffles before ONE set of
// atomics per (Gaussian, tile, warp), into per-Gaussian sums that the chain-rule kernel turns into parameter
// gradients:   gacc row = {dR, dG, dB, dZ | dOp, Sx, Sy, Sxx | Sxy, Syy, -, -},  S.. = sum dL/dpower * {dx, dy, dx^2, ..}.
#include <cuda_fp16.h>

#include "color.cuh"
#include "common.cuh"

namespace b2s {

constexpr float EXT_ALPHA_MAX = 0.999999f;   // oracle/ext_oracle.py ALPHA_MAX
constexpr float EXT_T_STOP = 1e-5f;          // a pixel stops compositing once its transmittance is below this
constexpr float LOG2E_F = 1.4426950408889634f;
constexpr int EXT_CHUNK = 128;               // Gaussians staged in shared memory per step

struct ExtGeom {          // everything the forward needs from the projection of one Gaussian
  float px, py, zabs, zcam;
  float A, B, C;          // conic
  float sx, sy;           // bbox sigmas
  int xmin, ymin, xmax, ymax;
  bool ok;
};

__device__ __forceinline__ void quat_rot(const float* q, float (&R)[3][3], float (&qn)[4], float& inv_norm) {
  const float nrm = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  inv_norm = 1.0f / (nrm + 1e-12f);
  const float w = q[0] * inv_norm, x = q[1] * inv_norm, y = q[2] * inv_norm, z = q[3] * inv_norm;
  qn[0] = w; qn[1] = x; qn[2] = y; qn[3] = z;
  R[0][0] = 1.f - 2.f * (y * y + z * z); R[0][1] = 2.f * (x * y - w * z); R[0][2] = 2.f * (x * z + w * y);
  R[1][0] = 2.f * (x * y + w * z); R[1][1] = 1.f - 2.f * (x * x + z * z); R[1][2] = 2.f * (y * z - w * x);
  R[2][0] = 2.f * (x * z - w * y); R[2][1] = 2.f * (y * z + w * x); R[2][2] = 1.f - 2.f * (x * x + y * y);
}

// Intermediates of the EWA projection kept for the chain rule.
struct EwaMats {
  float J[2][3], JW[2][3], T[2][3], M[3][3], R[3][3], qn[4], inv_norm;
  float N0[3], N1[3], iw2, ws;
  float c11, c12, c22;
};

// cam = V [m,1], clip = P cam (same op order as project_gaussian); s = |activated scales|, q = raw quaternion
__device__ __forceinline__ void ewa_project(const ViewParams& vp, const float* cam, const float* clip, const float* s,
                                            const float* q, float dilation, EwaMats& E) {
  const float w = clip[3];
  E.ws = (fabsf(w) < 1e-8f) ? 1.0f : w;
  E.iw2 = 1.0f / (E.ws * E.ws);
  const float a = 0.5f * vp.wm1, b = -0.5f * vp.hm1;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    E.N0[j] = vp.proj[j] * w - clip[0] * vp.proj[12 + j];
    E.N1[j] = vp.proj[4 + j] * w - clip[1] * vp.proj[12 + j];
    E.J[0][j] = a * E.N0[j] * E.iw2;
    E.J[1][j] = b * E.N1[j] * E.iw2;
  }
  quat_rot(q, E.R, E.qn, E.inv_norm);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) E.M[i][k] = E.R[i][k] * s[k];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c)
      E.JW[r][c] = E.J[r][0] * vp.view[c] + E.J[r][1] * vp.view[4 + c] + E.J[r][2] * vp.view[8 + c];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int k = 0; k < 3; ++k) E.T[r][k] = E.JW[r][0] * E.M[0][k] + E.JW[r][1] * E.M[1][k] + E.JW[r][2] * E.M[2][k];
  E.c11 = E.T[0][0] * E.T[0][0] + E.T[0][1] * E.T[0][1] + E.T[0][2] * E.T[0][2] + dilation;
  E.c12 = E.T[0][0] * E.T[1][0] + E.T[0][1] * E.T[1][1] + E.T[0][2] * E.T[1][2];
  E.c22 = E.T[1][0] * E.T[1][0] + E.T[1][1] * E.T[1][1] + E.T[1][2] * E.T[1][2] + dilation;
}

__device__ __forceinline__ void cam_clip(const ViewParams& vp, float mx, float my, float mz, float* cam, float* clip) {
#pragma unroll
  for (int q = 0; q < 4; ++q) cam[q] = dot4_lr(vp.view + 4 * q, mx, my, mz, 1.0f);
#pragma unroll
  for (int q = 0; q < 4; ++q) clip[q] = dot4_lr(vp.proj + 4 * q, cam[0], cam[1], cam[2], cam[3]);
}

// bbox of the k-sigma footprint, as project_gaussian computes it (same rounding order)
__device__ __forceinline__ bool ext_bbox(const ViewParams& vp, float px, float py, float sx, float sy, int& xmin, int& ymin,
                                         int& xmax, int& ymax) {
  const float rx = __fmul_rn(vp.k, sx), ry = __fmul_rn(vp.k, sy);
  const float lox = floorf(__fsub_rn(px, rx)), hix = ceilf(__fadd_rn(px, rx));
  const float loy = floorf(__fsub_rn(py, ry)), hiy = ceilf(__fadd_rn(py, ry));
  const bool ok = (hix >= 0.0f) && (lox <= vp.wm1) && (hiy >= 0.0f) && (loy <= vp.hm1);
  if (ok) {
    xmin = (int)fmaxf(lox, 0.0f); xmax = (int)fminf(hix, vp.wm1);
    ymin = (int)fmaxf(loy, 0.0f); ymax = (int)fminf(hiy, vp.hm1);
  } else {
    xmin = ymin = 0; xmax = ymax = -1;
  }
  return ok;
}

// ---- forward, per Gaussian ------------------------------------------------------------------------------------
// rotations == nullptr: the reference's axis-aligned sigmas (torch_renderer.py:147-150) as a conic with B = 0.
template <int K>
__global__ void __launch_bounds__(PRE_BLOCK)
ext_preprocess_kernel(const ViewParams vp, const float* __restrict__ means, const float* __restrict__ scales,
                      const float* __restrict__ rotations, const float* __restrict__ colors,
                      const float* __restrict__ opac, int n, float dilation, float4* __restrict__ rec,
                      uint8_t* __restrict__ cmask_out, uint2* __restrict__ rect, unsigned long long* __restrict__ tmask,
                      uint32_t* __restrict__ dbits, int* __restrict__ cnt, long long* __restrict__ bsum) {
  const int i = blockIdx.x * PRE_BLOCK + threadIdx.x;
  int my_cnt = 0;
  if (i < n) {
    const float mx = __ldg(means + 3 * (size_t)i), my = __ldg(means + 3 * (size_t)i + 1),
                mz = __ldg(means + 3 * (size_t)i + 2);
    const float s0 = act_scale(vp, __ldg(scales + 3 * (size_t)i)), s1 = act_scale(vp, __ldg(scales + 3 * (size_t)i + 1));
    const float op = act_opac(vp, __ldg(opac + i));
    const Proj pr = project_gaussian(vp, mx, my, mz, s0, s1, op);
    ExtGeom g;
    g.px = pr.px; g.py = pr.py; g.zabs = pr.zabs; g.zcam = pr.zcam;
    if (rotations == nullptr) {
      g.ok = pr.ok;
      g.sx = pr.sx; g.sy = pr.sy;
      g.A = 1.0f / (pr.sx * pr.sx); g.B = 0.0f; g.C = 1.0f / (pr.sy * pr.sy);
      g.xmin = pr.xmin; g.xmax = pr.xmax; g.ymin = pr.ymin; g.ymax = pr.ymax;
    } else {
      float cam[4], clip[4];
      cam_clip(vp, mx, my, mz, cam, clip);
      const float s[3] = {fabsf(s0), fabsf(s1), fabsf(act_scale(vp, __ldg(scales + 3 * (size_t)i + 2)))};
      const float q[4] = {__ldg(rotations + 4 * (size_t)i), __ldg(rotations + 4 * (size_t)i + 1),
                          __ldg(rotations + 4 * (size_t)i + 2), __ldg(rotations + 4 * (size_t)i + 3)};
      EwaMats E;
      ewa_project(vp, cam, clip, s, q, dilation, E);
      const float idet = 1.0f / (E.c11 * E.c22 - E.c12 * E.c12);
      g.A = E.c22 * idet; g.B = -E.c12 * idet; g.C = E.c11 * idet;
      g.sx = sqrtf(E.c11); g.sy = sqrtf(E.c22);
      g.ok = pr.valid && ext_bbox(vp, g.px, g.py, g.sx, g.sy, g.xmin, g.ymin, g.xmax, g.ymax);
    }
    uint2 rc = make_uint2(1u, 0u);
    unsigned long long tm = 0ull;
    if (g.ok) {
      const int tx0 = g.xmin / TILE, tx1 = g.xmax / TILE, ty0 = g.ymin / TILE, ty1 = g.ymax / TILE;
      const int w = tx1 - tx0 + 1, h = ty1 - ty0 + 1;
      my_cnt = w * h;
      if (w <= 8 && h <= 8) {   // explicit tile mask of small rects: every tile of the rect (no elliptical culling here)
        const unsigned long long rowm = (1ull << w) - 1ull;
        for (int r = 0; r < h; ++r) tm |= rowm << (r * w);
      }
      rc = make_uint2((uint32_t)tx0 | ((uint32_t)ty0 << 16), (uint32_t)tx1 | ((uint32_t)ty1 << 16));
    }
    rect[i] = rc;
    tmask[i] = tm;
    dbits[i] = depth_bits(g.zcam);
    cnt[i] = my_cnt;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = make_float4(0.f, 0.f, 0.f, __int_as_float(0)), c = make_float4(0.f, 0.f, 0.f, __int_as_float(0));
    int cm = 0;
    if (g.ok) {
      float craw[3], dir[3], rinv;
      float coef[K * 3];
      load_coeffs<K>(colors, i, coef);
      eval_color<K>(vp, coef, mx, my, mz, craw, dir, &rinv);
      a = make_float4(g.px, g.py, -0.5f * LOG2E_F * g.A, -0.5f * LOG2E_F * g.C);
      b = make_float4(-LOG2E_F * g.B, op, g.zabs, __int_as_float(g.xmin | (g.xmax << 16)));
      c = make_float4(fminf(fmaxf(craw[0], 0.0f), 1.0f), fminf(fmaxf(craw[1], 0.0f), 1.0f), fminf(fmaxf(craw[2], 0.0f), 1.0f),
                      __int_as_float(g.ymin | (g.ymax << 16)));
      cm = (craw[0] >= 0.0f && craw[0] <= 1.0f ? 1 : 0) | (craw[1] >= 0.0f && craw[1] <= 1.0f ? 2 : 0) |
           (craw[2] >= 0.0f && craw[2] <= 1.0f ? 4 : 0);
    }
    rec[3 * (size_t)i] = a;
    rec[3 * (size_t)i + 1] = b;
    rec[3 * (size_t)i + 2] = c;
    cmask_out[i] = (uint8_t)cm;
  }
  __shared__ int wsum[PRE_BLOCK / 32];
  int s = my_cnt;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t = 0;
#pragma unroll
    for (int q = 0; q < PRE_BLOCK / 32; ++q) t += wsum[q];
    bsum[blockIdx.x] = t;
  }
}

int launch_ext_preprocess(const ViewParams& vp, const float* means, const float* scales, const float* rotations,
                          const float* colors, const float* opac, int n, float dilation, float4* rec, uint8_t* cmask,
                          uint2* rect, unsigned long long* tmask, uint32_t* dbits, int* cnt, long long* bsum,
                          cudaStream_t st) {
  if (n <= 0) return B2S_OK;
  const int blocks = (n + PRE_BLOCK - 1) / PRE_BLOCK;
#define B2S_EXTPRE(KK) ext_preprocess_kernel<KK><<<blocks, PRE_BLOCK, 0, st>>>(vp, means, scales, rotations, colors, opac, n, dilation, rec, cmask, rect, tmask, dbits, cnt, bsum)
  switch (vp.sh) {
    case 1: B2S_EXTPRE(1); break;
    case 4: B2S_EXTPRE(4); break;
    case 9: B2S_EXTPRE(9); break;
    case 16: B2S_EXTPRE(16); break;
    default: set_error("sh_coeffs must be 1, 4, 9 or 16 (got %d)", vp.sh); return B2S_ERR_INVALID;
  }
#undef B2S_EXTPRE
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

// ---- blend, per pixel ---------------------------------------------------------------------------------------------
struct ExtStage {
  float4 a[EXT_CHUNK], b[EXT_CHUNK], c[EXT_CHUNK];
};

__device__ __forceinline__ void ext_stage_chunk(ExtStage& s, const float4* __restrict__ rec, const int* __restrict__ vals,
                                                int first, int cnt) {
  for (int j = threadIdx.x; j < cnt; j += TILE_PIX) {
    const int id = __ldg(vals + first + j);
    s.a[j] = __ldg(rec + 3 * (size_t)id);
    s.b[j] = __ldg(rec + 3 * (size_t)id + 1);
    s.c[j] = __ldg(rec + 3 * (size_t)id + 2);
  }
}

// weight of staged Gaussian j at pixel centre (fx, fy) / integer pixel (x, y).  OVER: alpha with the reference's rules
// (pixel bbox, a < 1e-5 dropped, capped); returns 0 for "no contribution".  G = 2^power (without opacity).
template <bool OVER>
__device__ __forceinline__ float ext_weight(const ExtStage& s, int j, float fx, float fy, int x, int y, float& dx, float& dy,
                                            float& G, bool& capped) {
  const float4 a = s.a[j], b = s.b[j];
  dx = fx - a.x;
  dy = fy - a.y;
  const float p2 = fmaf(a.z * dx, dx, fmaf(a.w * dy, dy, b.x * dx * dy));
  G = ex2_approx(p2);
  capped = false;
  float w = b.y * G;
  if (OVER) {
    const int bx = __float_as_int(b.w), by = __float_as_int(s.c[j].w);
    const bool in = (x >= (bx & 0xffff)) && (x <= (bx >> 16)) && (y >= (by & 0xffff)) && (y <= (by >> 16));
    if (!in || !(w >= 1e-5f)) return 0.0f;
    if (w > EXT_ALPHA_MAX) { w = EXT_ALPHA_MAX; capped = true; }
  }
  return w;
}

// acc planes (H*W each): WSUM {Cr, Cg, Cb, W, D};  OVER {Cr, Cg, Cb, T_end, number of list entries consumed}
template <bool OVER>
__global__ void __launch_bounds__(TILE_PIX)
blend_ext_fwd_kernel(const ViewParams vp, const float4* __restrict__ rec, const int* __restrict__ vals,
                     const int2* __restrict__ ranges, float* __restrict__ out_rgb, float* __restrict__ out_alpha,
                     float* __restrict__ out_depth, float* __restrict__ acc) {
  __shared__ ExtStage st;
  const int tile = blockIdx.x;
  const int tx = tile % vp.tiles_x, ty = tile / vp.tiles_x;
  const int col = threadIdx.x & (TILE - 1), row = threadIdx.x / TILE;
  const int x = tx * TILE + col, y = ty * TILE + row;
  const bool inside = x < vp.width && y < vp.height;
  const float fx = (float)x + 0.5f, fy = (float)y + 0.5f;
  const int2 rg = ranges[tile];
  const int L = rg.y - rg.x;
  float c0 = 0.f, c1 = 0.f, c2 = 0.f, W = 0.f, D = 0.f, T = 1.0f;
  int used = 0;
  bool done = !inside;
  for (int base = 0; base < L; base += EXT_CHUNK) {
    const int cnt = min(EXT_CHUNK, L - base);
    __syncthreads();
    ext_stage_chunk(st, rec, vals, rg.x + base, cnt);
    __syncthreads();
    if (!done) {
      for (int j = 0; j < cnt; ++j) {
        float dx, dy, G;
        bool capped;
        const float w = ext_weight<OVER>(st, j, fx, fy, x, y, dx, dy, G, capped);
        if (OVER) {
          if (w > 0.0f) {
            const float4 cc = st.c[j];
            const float contrib = T * w;
            c0 = fmaf(contrib, cc.x, c0); c1 = fmaf(contrib, cc.y, c1); c2 = fmaf(contrib, cc.z, c2);
            D = fmaf(contrib, st.b[j].z, D);
            T *= (1.0f - w);
          }
          used = base + j + 1;
          if (T < EXT_T_STOP) { done = true; break; }
        } else {
          const float4 cc = st.c[j];
          c0 = fmaf(w, cc.x, c0); c1 = fmaf(w, cc.y, c1); c2 = fmaf(w, cc.z, c2);
          W += w;
          D = fmaf(w, st.b[j].z, D);
        }
      }
    }
    if (OVER && __syncthreads_and(done)) break;
  }
  if (!inside) return;
  const size_t hw = (size_t)vp.width * vp.height, p = (size_t)y * vp.width + x;
  const float b0 = view_bg(vp, 0), b1 = view_bg(vp, 1), b2 = view_bg(vp, 2);
  float r, g, b, alpha, depth;
  if (OVER) {
    r = fmaf(T, b0, c0); g = fmaf(T, b1, c1); b = fmaf(T, b2, c2);
    alpha = 1.0f - T;
    depth = D;
    acc[p] = c0; acc[hw + p] = c1; acc[2 * hw + p] = c2; acc[3 * hw + p] = T; acc[4 * hw + p] = __int_as_float(used);
  } else {
    const float inv = 1.0f / (1.0f + W);
    r = (b0 + c0) * inv; g = (b1 + c1) * inv; b = (b2 + c2) * inv;
    alpha = W * inv;
    depth = fmaxf(D / (W + 1e-6f), 0.0f);
    acc[p] = c0; acc[hw + p] = c1; acc[2 * hw + p] = c2; acc[3 * hw + p] = W; acc[4 * hw + p] = D;
  }
  if (out_rgb != nullptr) {
    out_rgb[3 * p] = fminf(fmaxf(r, 0.0f), 1.0f);
    out_rgb[3 * p + 1] = fminf(fmaxf(g, 0.0f), 1.0f);
    out_rgb[3 * p + 2] = fminf(fmaxf(b, 0.0f), 1.0f);
  }
  if (out_alpha != nullptr) out_alpha[p] = fminf(fmaxf(alpha, 0.0f), 1.0f);
  if (out_depth != nullptr) out_depth[p] = depth;
}

int launch_blend_ext_fwd(const ViewParams& vp, int over, const float4* rec, const int* vals, const int2* ranges,
                         float* out_rgb, float* out_alpha, float* out_depth, float* acc, cudaStream_t st) {
  if (over) blend_ext_fwd_kernel<true><<<vp.n_tiles, TILE_PIX, 0, st>>>(vp, rec, vals, ranges, out_rgb, out_alpha, out_depth, acc);
  else blend_ext_fwd_kernel<false><<<vp.n_tiles, TILE_PIX, 0, st>>>(vp, rec, vals, ranges, out_rgb, out_alpha, out_depth, acc);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

__global__ void __launch_bounds__(256) ext_gacc_zero_kernel(float4* __restrict__ gacc, int n3) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < n3) gacc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Backward of the blend.  WSUM: every list entry, any order.  OVER: back to front from the last entry the pixel
// consumed, T_i recovered by dividing by (1 - a_i) (a_i <= EXT_ALPHA_MAX), the colour behind by the recursion
// B <- a v + (1 - a) B -- no subtraction of nearly equal sums.
template <bool OVER>
__global__ void __launch_bounds__(TILE_PIX)
blend_ext_bwd_kernel(const ViewParams vp, const float4* __restrict__ rec, const int* __restrict__ vals,
                     const int2* __restrict__ ranges, const float* __restrict__ acc, const float* __restrict__ g_rgb,
                     const float* __restrict__ g_alpha, const float* __restrict__ g_depth, float* __restrict__ gacc) {
  __shared__ ExtStage st;
  __shared__ int s_max;
  const int tile = blockIdx.x;
  const int tx = tile % vp.tiles_x, ty = tile / vp.tiles_x;
  const int col = threadIdx.x & (TILE - 1), row = threadIdx.x / TILE;
  const int x = tx * TILE + col, y = ty * TILE + row;
  const bool inside = x < vp.width && y < vp.height;
  const float fx = (float)x + 0.5f, fy = (float)y + 0.5f;
  const int lane = threadIdx.x & 31;
  const int2 rg = ranges[tile];
  const int L = rg.y - rg.x;
  const size_t hw = (size_t)vp.width * vp.height, p = (size_t)y * vp.width + x;
  // upstream gradients on the accumulators: gam[0..2] colour planes, gam[3] depth plane, gW (WSUM: on W; OVER: on A = 1 - T)
  float gam[4] = {0.f, 0.f, 0.f, 0.f}, gW = 0.f;
  float T = 1.0f, Bc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  int used = 0;
  if (inside) {
    const float a0 = acc[p], a1 = acc[hw + p], a2 = acc[2 * hw + p], a3 = acc[3 * hw + p], a4 = acc[4 * hw + p];
    const float bg[3] = {view_bg(vp, 0), view_bg(vp, 1), view_bg(vp, 2)};
    const float gr[3] = {g_rgb ? g_rgb[3 * p] : 0.f, g_rgb ? g_rgb[3 * p + 1] : 0.f, g_rgb ? g_rgb[3 * p + 2] : 0.f};
    const float ga = g_alpha ? g_alpha[p] : 0.f, gd = g_depth ? g_depth[p] : 0.f;
    const float cs[3] = {a0, a1, a2};
    if (OVER) {
      T = a3;
      used = __float_as_int(a4);
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const float o = fmaf(T, bg[q], cs[q]);
        gam[q] = (o >= 0.0f && o <= 1.0f) ? gr[q] : 0.0f;
        Bc[q] = bg[q];
      }
      gam[3] = gd;
      const float al = 1.0f - T;
      gW = (al >= 0.0f && al <= 1.0f) ? ga : 0.0f;
    } else {
      const float W = a3, D = a4, inv = 1.0f / (1.0f + W);
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const float o = (bg[q] + cs[q]) * inv;
        const float gq = (o >= 0.0f && o <= 1.0f) ? gr[q] : 0.0f;
        gam[q] = gq * inv;
        gW -= gq * o * inv;
      }
      const float al = W * inv;
      if (al >= 0.0f && al <= 1.0f) gW += ga * inv * inv;
      const float wd = W + 1e-6f;
      if (D / wd >= 0.0f) { gam[3] = gd / wd; gW -= gd * D / (wd * wd); }
      used = L;
    }
  }
  // the tile walks as far as its deepest pixel went
  if (threadIdx.x == 0) s_max = 0;
  __syncthreads();
  if (OVER) { if (used > 0) atomicMax(&s_max, used); } else if (threadIdx.x == 0) s_max = L;
  __syncthreads();
  const int Lmax = s_max;
  const int nchunks = (Lmax + EXT_CHUNK - 1) / EXT_CHUNK;
  for (int ci = 0; ci < nchunks; ++ci) {
    const int c = OVER ? nchunks - 1 - ci : ci;
    const int base = c * EXT_CHUNK;
    const int cnt = min(EXT_CHUNK, Lmax - base);
    __syncthreads();
    ext_stage_chunk(st, rec, vals, rg.x + base, cnt);
    __syncthreads();
    for (int jj = 0; jj < cnt; ++jj) {
      const int j = OVER ? cnt - 1 - jj : jj;
      float v[10] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      bool hit = false;
      if (inside && base + j < used) {
        float dx, dy, G;
        bool capped;
        const float w = ext_weight<OVER>(st, j, fx, fy, x, y, dx, dy, G, capped);
        const float4 cc = st.c[j];
        const float z = st.b[j].z;
        if (OVER) {
          if (w > 0.0f) {
            const float om = 1.0f - w;
            const float Ti = T / om;
            const float dLda = Ti * (gam[0] * (cc.x - Bc[0]) + gam[1] * (cc.y - Bc[1]) + gam[2] * (cc.z - Bc[2]) +
                                     gam[3] * (z - Bc[3]) + gW * (1.0f - Bc[4]));
            const float tw = Ti * w;
            v[0] = gam[0] * tw; v[1] = gam[1] * tw; v[2] = gam[2] * tw; v[3] = gam[3] * tw;
            if (!capped) {
              v[4] = dLda * G;
              const float dLdp = dLda * w;
              v[5] = dLdp * dx; v[6] = dLdp * dy; v[7] = dLdp * dx * dx; v[8] = dLdp * dx * dy; v[9] = dLdp * dy * dy;
            }
            Bc[0] = fmaf(w, cc.x, om * Bc[0]); Bc[1] = fmaf(w, cc.y, om * Bc[1]); Bc[2] = fmaf(w, cc.z, om * Bc[2]);
            Bc[3] = fmaf(w, z, om * Bc[3]); Bc[4] = fmaf(w, 1.0f, om * Bc[4]);
            T = Ti;
            hit = true;
          }
        } else {
          const float dLdw = gam[0] * cc.x + gam[1] * cc.y + gam[2] * cc.z + gam[3] * z + gW;
          v[0] = gam[0] * w; v[1] = gam[1] * w; v[2] = gam[2] * w; v[3] = gam[3] * w;
          v[4] = dLdw * G;
          const float dLdp = dLdw * w;
          v[5] = dLdp * dx; v[6] = dLdp * dy; v[7] = dLdp * dx * dx; v[8] = dLdp * dx * dy; v[9] = dLdp * dy * dy;
          // a pixel further than ~5.7 sigma from the Gaussian (G < 1e-7) moves no gradient at the 1e-3 bar; without this
          // every warp of the tile reduces every list entry (the weighted sum has no pixel bbox) -- measured 6.7 ms
          // against 1.1 ms for the bbox-masked "over" blend on the same scene.  Lanes below the bar still add their
          // (tiny, exact) terms whenever another lane of the warp is above it.
          hit = G > 1e-7f;
        }
      }
      if (__any_sync(0xffffffffu, hit)) {
#pragma unroll
        for (int q = 0; q < 10; ++q) v[q] = warp_sum(v[q]);
        if (lane == 0) {
          float* dst = gacc + (size_t)__ldg(vals + rg.x + base + j) * GACC_F;
#pragma unroll
          for (int q = 0; q < 10; ++q)
            if (v[q] != 0.0f) atomicAdd(dst + q, v[q]);
        }
      }
    }
  }
}

int launch_blend_ext_bwd(const ViewParams& vp, int over, const float4* rec, const int* vals, const int2* ranges,
                         const float* acc, const float* g_rgb, const float* g_alpha, const float* g_depth, float* gacc,
                         int n, cudaStream_t st) {
  const int n3 = n * (GACC_F / 4);
  ext_gacc_zero_kernel<<<(n3 + 255) / 256, 256, 0, st>>>((float4*)gacc, n3);
  B2S_LAUNCH_CHECK();
  if (over) blend_ext_bwd_kernel<true><<<vp.n_tiles, TILE_PIX, 0, st>>>(vp, rec, vals, ranges, acc, g_rgb, g_alpha, g_depth, gacc);
  else blend_ext_bwd_kernel<false><<<vp.n_tiles, TILE_PIX, 0, st>>>(vp, rec, vals, ranges, acc, g_rgb, g_alpha, g_depth, gacc);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

// ---- chain rule, per Gaussian: gacc sums -> parameter gradients ---------------------------------------------------
template <int K>
__global__ void __launch_bounds__(PRE_BLOCK)
ext_bwd_kernel(const ViewParams vp, const float* __restrict__ means, const float* __restrict__ scales,
               const float* __restrict__ rotations, const float* __restrict__ colors, const float* __restrict__ opac,
               int n, float dilation, const float4* __restrict__ gacc, const uint8_t* __restrict__ cmask,
               float* __restrict__ g_means, float* __restrict__ g_scales, float* __restrict__ g_rot,
               float* __restrict__ g_colors, float* __restrict__ g_opac) {
  const int i = blockIdx.x * PRE_BLOCK + threadIdx.x;
  if (i >= n) return;
  const float mx = __ldg(means + 3 * (size_t)i), my = __ldg(means + 3 * (size_t)i + 1), mz = __ldg(means + 3 * (size_t)i + 2);
  const float raw_s[3] = {__ldg(scales + 3 * (size_t)i), __ldg(scales + 3 * (size_t)i + 1), __ldg(scales + 3 * (size_t)i + 2)};
  const float raw_op = __ldg(opac + i);
  const float sa[3] = {act_scale(vp, raw_s[0]), act_scale(vp, raw_s[1]), act_scale(vp, raw_s[2])};
  const float op = act_opac(vp, raw_op);
  const float4 g0 = __ldg(gacc + 3 * (size_t)i), g1 = __ldg(gacc + 3 * (size_t)i + 1), g2 = __ldg(gacc + 3 * (size_t)i + 2);
  const float dC[3] = {g0.x, g0.y, g0.z};
  float dZ = g0.w;
  const float dOp = g1.x, Sx = g1.y, Sy = g1.z, Sxx = g1.w, Sxy = g2.x, Syy = g2.y;
  const Proj pr = project_gaussian(vp, mx, my, mz, sa[0], sa[1], op);
  float gm[3] = {0.f, 0.f, 0.f}, gs[3] = {0.f, 0.f, 0.f}, gq[4] = {0.f, 0.f, 0.f, 0.f}, gop = 0.f;
  float gcoef[K * 3];
#pragma unroll
  for (int q = 0; q < K * 3; ++q) gcoef[q] = 0.0f;
  // a culled Gaussian has an all-zero gacc row; valid-but-off-screen EWA footprints are covered by the same rule
  bool ok = rotations == nullptr ? pr.ok : pr.valid;
  if (ok) {
    float go = dOp;
    if (vp.act & B2S_ACT_OPACITY_SIGMOID) go *= op * (1.0f - op);
    gop = go;
    float dcam[4] = {0.f, 0.f, 0.f, 0.f};
    float dpx, dpy;
    if (rotations == nullptr) {
      const float isx2 = 1.0f / (pr.sx * pr.sx), isy2 = 1.0f / (pr.sy * pr.sy);
      dpx = Sx * isx2;
      dpy = Sy * isy2;
      const float dsx = Sxx * isx2 / pr.sx, dsy = Syy * isy2 / pr.sy;
      if (pr.ax >= 1.0f) {
        const float sgn = (float)((sa[0] > 0.f) - (sa[0] < 0.f));
        float t = dsx * sgn * (0.5f * vp.wf * vp.fx / pr.zabs);
        if (vp.act & B2S_ACT_SCALES_SOFTPLUS) t *= sigmoidf_acc(raw_s[0]);
        gs[0] = t;
        dZ -= dsx * pr.ax / pr.zabs;
      }
      if (pr.ay >= 1.0f) {
        const float sgn = (float)((sa[1] > 0.f) - (sa[1] < 0.f));
        float t = dsy * sgn * (0.5f * vp.hf * vp.fy / pr.zabs);
        if (vp.act & B2S_ACT_SCALES_SOFTPLUS) t *= sigmoidf_acc(raw_s[1]);
        gs[1] = t;
        dZ -= dsy * pr.ay / pr.zabs;
      }
    } else {
      float cam[4], clip[4];
      cam_clip(vp, mx, my, mz, cam, clip);
      const float s[3] = {fabsf(sa[0]), fabsf(sa[1]), fabsf(sa[2])};
      const float q[4] = {__ldg(rotations + 4 * (size_t)i), __ldg(rotations + 4 * (size_t)i + 1),
                          __ldg(rotations + 4 * (size_t)i + 2), __ldg(rotations + 4 * (size_t)i + 3)};
      EwaMats E;
      ewa_project(vp, cam, clip, s, q, dilation, E);
      const float idet = 1.0f / (E.c11 * E.c22 - E.c12 * E.c12);
      const float A = E.c22 * idet, B = -E.c12 * idet, C = E.c11 * idet;
      dpx = A * Sx + B * Sy;
      dpy = B * Sx + C * Sy;
      // conic -> covariance:  G_M = -Q G_Q Q,  G_Q = [[gA, gB/2], [gB/2, gC]]
      const float gA = -0.5f * Sxx, gBh = -0.5f * Sxy, gC = -0.5f * Syy;
      const float X00 = A * gA + B * gBh, X01 = A * gBh + B * gC, X10 = B * gA + C * gBh, X11 = B * gBh + C * gC;
      const float G00 = -(X00 * A + X01 * B), G01 = -(X00 * B + X01 * C), G10 = -(X10 * A + X11 * B), G11 = -(X10 * B + X11 * C);
      const float Gs = 0.5f * (G01 + G10);
      float dT[2][3], dJW[2][3], dM[3][3], dJ[2][3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        dT[0][k] = 2.0f * (G00 * E.T[0][k] + Gs * E.T[1][k]);
        dT[1][k] = 2.0f * (Gs * E.T[0][k] + G11 * E.T[1][k]);
      }
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) dJW[r][c] = dT[r][0] * E.M[c][0] + dT[r][1] * E.M[c][1] + dT[r][2] * E.M[c][2];
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int k = 0; k < 3; ++k) dM[c][k] = E.JW[0][c] * dT[0][k] + E.JW[1][c] * dT[1][k];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) dJ[r][c] = dJW[r][0] * vp.view[4 * c] + dJW[r][1] * vp.view[4 * c + 1] + dJW[r][2] * vp.view[4 * c + 2];
      // scales and rotation
      float dR[3][3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float t = E.R[0][k] * dM[0][k] + E.R[1][k] * dM[1][k] + E.R[2][k] * dM[2][k];
        t *= (float)((sa[k] > 0.f) - (sa[k] < 0.f));
        if (vp.act & B2S_ACT_SCALES_SOFTPLUS) t *= sigmoidf_acc(raw_s[k]);
        gs[k] = t;
#pragma unroll
        for (int r = 0; r < 3; ++r) dR[r][k] = dM[r][k] * s[k];
      }
      const float w = E.qn[0], x = E.qn[1], y = E.qn[2], z = E.qn[3];
      float dq[4];
      dq[0] = 2.0f * (-z * dR[0][1] + y * dR[0][2] + z * dR[1][0] - x * dR[1][2] - y * dR[2][0] + x * dR[2][1]);
      dq[1] = 2.0f * (y * dR[0][1] + z * dR[0][2] + y * dR[1][0] - 2.0f * x * dR[1][1] - w * dR[1][2] + z * dR[2][0] + w * dR[2][1] - 2.0f * x * dR[2][2]);
      dq[2] = 2.0f * (-2.0f * y * dR[0][0] + x * dR[0][1] + w * dR[0][2] + x * dR[1][0] + z * dR[1][2] - w * dR[2][0] + z * dR[2][1] - 2.0f * y * dR[2][2]);
      dq[3] = 2.0f * (-2.0f * z * dR[0][0] - w * dR[0][1] + x * dR[0][2] + w * dR[1][0] - 2.0f * z * dR[1][1] + y * dR[1][2] + x * dR[2][0] + y * dR[2][1]);
      const float dot = w * dq[0] + x * dq[1] + y * dq[2] + z * dq[3];
#pragma unroll
      for (int k = 0; k < 4; ++k) gq[k] = (dq[k] - E.qn[k] * dot) * E.inv_norm;
      // Jacobian -> camera-space position
      if (fabsf(clip[3]) >= 1e-8f) {
        const float a = 0.5f * vp.wm1, b = -0.5f * vp.hm1;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float p3k = vp.proj[12 + k];
          float t = 0.0f;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const float p3j = vp.proj[12 + j];
            t += dJ[0][j] * a * ((vp.proj[j] * p3k - vp.proj[k] * p3j) * E.iw2 - 2.0f * E.N0[j] * p3k * E.iw2 / E.ws);
            t += dJ[1][j] * b * ((vp.proj[4 + j] * p3k - vp.proj[4 + k] * p3j) * E.iw2 - 2.0f * E.N1[j] * p3k * E.iw2 / E.ws);
          }
          dcam[k] += t;
        }
      }
    }
    // zabs = max(|cam.z|, 1e-6)
    if (fabsf(pr.zcam) >= 1e-6f) dcam[2] += dZ * ((pr.zcam > 0.f) ? 1.0f : -1.0f);
    // px,py -> ndc -> clip -> cam
    const float dnx = dpx * 0.5f * vp.wm1, dny = -dpy * 0.5f * vp.hm1;
    float dclip[4];
    dclip[0] = dnx / pr.wsafe;
    dclip[1] = dny / pr.wsafe;
    dclip[2] = 0.0f;
    dclip[3] = (fabsf(pr.w) < 1e-8f) ? 0.0f : -(dnx * pr.ndcx + dny * pr.ndcy) / pr.wsafe;
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int r = 0; r < 4; ++r) dcam[c] = fmaf(vp.proj[4 * r + c], dclip[r], dcam[c]);
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int r = 0; r < 4; ++r) gm[c] = fmaf(vp.view[4 * r + c], dcam[r], gm[c]);
    // colour
    const int cm = cmask[i];
    const float dc0 = (cm & 1) ? dC[0] : 0.0f, dc1 = (cm & 2) ? dC[1] : 0.0f, dc2 = (cm & 4) ? dC[2] : 0.0f;
    if constexpr (K == 1) {
      float craw[3] = {__ldg(colors + 3 * (size_t)i), __ldg(colors + 3 * (size_t)i + 1), __ldg(colors + 3 * (size_t)i + 2)};
      const float dcs[3] = {dc0, dc1, dc2};
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        if (vp.act & B2S_ACT_COLORS_SIGMOID) {
          const float sg = sigmoidf_acc(craw[q]);
          gcoef[q] = dcs[q] * sg * (1.0f - sg);
        } else {
          gcoef[q] = dcs[q];
        }
      }
    } else {
      float coef[K * 3];
      load_coeffs<K>(colors, i, coef);
      const float vx = vp.cam[0] - mx, vy = vp.cam[1] - my, vz = vp.cam[2] - mz;
      const float r = sqrtf(vx * vx + vy * vy + vz * vz);
      const float rinv = 1.0f / (r + 1e-8f);
      const float ddx = vx * rinv, ddy = vy * rinv, ddz = vz * rinv;
      float bs[16], bx[16], by[16], bz[16];
      sh_basis(ddx, ddy, ddz, K, bs);
      sh_basis_grad(ddx, ddy, ddz, K, bx, by, bz);
      float dd0 = 0.f, dd1 = 0.f, dd2 = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const float sdot = coef[3 * k] * dc0 + coef[3 * k + 1] * dc1 + coef[3 * k + 2] * dc2;
        gcoef[3 * k] = bs[k] * dc0; gcoef[3 * k + 1] = bs[k] * dc1; gcoef[3 * k + 2] = bs[k] * dc2;
        dd0 = fmaf(bx[k], sdot, dd0); dd1 = fmaf(by[k], sdot, dd1); dd2 = fmaf(bz[k], sdot, dd2);
      }
      const float vdd = vx * dd0 + vy * dd1 + vz * dd2;
      const float k2 = (r > 0.0f) ? vdd * rinv * rinv / r : 0.0f;
      gm[0] -= dd0 * rinv - vx * k2;
      gm[1] -= dd1 * rinv - vy * k2;
      gm[2] -= dd2 * rinv - vz * k2;
    }
  }
  g_means[3 * (size_t)i] = gm[0]; g_means[3 * (size_t)i + 1] = gm[1]; g_means[3 * (size_t)i + 2] = gm[2];
  g_scales[3 * (size_t)i] = gs[0]; g_scales[3 * (size_t)i + 1] = gs[1]; g_scales[3 * (size_t)i + 2] = gs[2];
  if (g_rot != nullptr) {
    g_rot[4 * (size_t)i] = gq[0]; g_rot[4 * (size_t)i + 1] = gq[1]; g_rot[4 * (size_t)i + 2] = gq[2]; g_rot[4 * (size_t)i + 3] = gq[3];
  }
  g_opac[i] = gop;
#pragma unroll
  for (int q = 0; q < K * 3; ++q) g_colors[(size_t)i * K * 3 + q] = gcoef[q];
}

int launch_ext_bwd(const ViewParams& vp, const float* means, const float* scales, const float* rotations,
                   const float* colors, const float* opac, int n, float dilation, const float* gacc, const uint8_t* cmask,
                   float* g_means, float* g_scales, float* g_rot, float* g_colors, float* g_opac, cudaStream_t st) {
  if (n <= 0) return B2S_OK;
  const int blocks = (n + PRE_BLOCK - 1) / PRE_BLOCK;
#define B2S_EXTBWD(KK) ext_bwd_kernel<KK><<<blocks, PRE_BLOCK, 0, st>>>(vp, means, scales, rotations, colors, opac, n, dilation, (const float4*)gacc, cmask, g_means, g_scales, g_rot, g_colors, g_opac)
  switch (vp.sh) {
    case 1: B2S_EXTBWD(1); break;
    case 4: B2S_EXTBWD(4); break;
    case 9: B2S_EXTBWD(9); break;
    case 16: B2S_EXTBWD(16); break;
    default: set_error("sh_coeffs must be 1, 4, 9 or 16 (got %d)", vp.sh); return B2S_ERR_INVALID;
  }
#undef B2S_EXTBWD
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

}  // namespace b2s
