// Per-Gaussian kernels: projection + sigma + colour/SH + bbox + tile count (forward) and the
// chain rule back to means / scales / opacities / colours (backward).
//
// Replaces: _project / sigma / _eval_colors of the reference
//   python/torch_renderer.py:57-106,143-150   and the per-thread prologue of
//   src/renderer.cu:41-84 (AoS stride-3 loads, no culling of work).
// HBM-bound: algorithmic bytes per Gaussian*view = 28 + 12*sh (read) + 48+16 (write).
#include <cuda_fp16.h>
#include <string.h>

#include "color.cuh"
#include "common.cuh"

namespace b2s {


// x in [0,1] as two fp16 numbers hi + lo (22 significant bits), packed {low half = hi, high half = lo}
__device__ __forceinline__ float split_f16_pair(float x) {
  const __half hi = __float2half_rn(x);
  const __half lo = __float2half_rn(x - __half2float(hi));
  return __uint_as_float((uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16));
}

// Blend record + clamp mask of one projected Gaussian (layout: see preprocess_kernel).
__device__ __forceinline__ void make_record(const ViewParams& vp, const Proj& pr, float op, const float* craw, float4& a,
                                            float4& b, float4& c, int& cm) {
  // The reference's Gaussians are axis aligned, so the weight is separable:
  //   w(x,y) = [op * exp(-dx^2/2sx^2)] * [exp(-dy^2/2sy^2)]  -- one x record, one y record.
  const bool lg = (vp.exact_bbox == 0);        // log-domain opacity unless the native exact mode
  c.x = fminf(fmaxf(craw[0], 0.0f), 1.0f);
  c.y = fminf(fmaxf(craw[1], 0.0f), 1.0f);
  c.z = fminf(fmaxf(craw[2], 0.0f), 1.0f);
  c.w = pr.zabs;
  a.x = pr.px;
  a.y = NEG_HALF_LOG2E / (pr.sx * pr.sx);
  a.z = lg ? log2f(op) : op;
  a.w = lg ? split_f16_pair(c.x) : __int_as_float(pr.xmin | (pr.xmax << 16));
  b.x = pr.py;
  b.y = NEG_HALF_LOG2E / (pr.sy * pr.sy);
  b.z = lg ? split_f16_pair(c.z) : 1.0f;
  b.w = lg ? split_f16_pair(c.y) : __int_as_float(pr.ymin | (pr.ymax << 16));
  cm = (craw[0] >= 0.0f && craw[0] <= 1.0f ? 1 : 0) | (craw[1] >= 0.0f && craw[1] <= 1.0f ? 2 : 0) |
       (craw[2] >= 0.0f && craw[2] <= 1.0f ? 4 : 0);
}

// Forward: one thread per Gaussian.  Writes the 48-byte blend record
//   rec[3i+0] = {px, qx, lop, red  as f16 hi|lo}   qx = -0.5*log2(e)/sx^2, lop = log2(op)
//   rec[3i+1] = {py, qy, blue as f16 hi|lo, green as f16 hi|lo}   so  w = 2^(qx dx^2 + lop) * 2^(qy dy^2)
//   rec[3i+2] = {r, g, b, zabs}
//   The f16 pairs (low half = fp16(c), high half = fp16(c - hi)) are the clamped colour pre-split for the
//   tensor-core forward, which forms its fp16 hi/lo B operands from them with packed half arithmetic.
//   exact_bbox mode (native styles) keeps  {px, qx, op, bbox x (min|max<<16)}, {py, qy, 1, bbox y}:
//   w = op*2^(qx dx^2) * 1*2^(qy dy^2), cut to the pixel bbox.
// cmask[i] (torch style): bit q = colour channel q is inside [0,1], i.e. clamp(0,1) passes its gradient
// (torch_renderer.py:144); consumed by the blend backward through gacc_init_kernel.
// plus the tile rect / depth bits / tile count consumed by the binning kernels, and the
// per-block sum of tile counts (first level of the exclusive scan).
template <int K>
__global__ void __launch_bounds__(PRE_BLOCK)
preprocess_kernel(const ViewParams vp, const float* __restrict__ means, const float* __restrict__ scales,
                  const float* __restrict__ colors, const float* __restrict__ opac, int n,
                  float4* __restrict__ rec, uint8_t* __restrict__ cmask_out, uint2* __restrict__ rect,
                  unsigned long long* __restrict__ tmask,
                  uint32_t* __restrict__ dbits, int* __restrict__ cnt, long long* __restrict__ bsum,
                  float* __restrict__ dbg, int* __restrict__ dbg_bbox) {
  const int i = blockIdx.x * PRE_BLOCK + threadIdx.x;
  int my_cnt = 0;
  if (i < n) {
    const float mx = __ldg(means + 3 * (size_t)i), my = __ldg(means + 3 * (size_t)i + 1),
                mz = __ldg(means + 3 * (size_t)i + 2);
    const float s0 = act_scale(vp, __ldg(scales + 3 * (size_t)i)), s1 = act_scale(vp, __ldg(scales + 3 * (size_t)i + 1));
    const float op = act_opac(vp, __ldg(opac + i));
    const Proj pr = project_gaussian(vp, mx, my, mz, s0, s1, op);
    uint2 rc = make_uint2(1u, 0u);   // empty tile rect (tx1 < tx0) for culled Gaussians
    unsigned long long tm = 0ull;
    if (pr.ok) {
      const int tx0 = pr.xmin / TILE, tx1 = pr.xmax / TILE, ty0 = pr.ymin / TILE, ty1 = pr.ymax / TILE;
      const int w = tx1 - tx0 + 1, h = ty1 - ty0 + 1;
      my_cnt = w * h;
      if (w <= 8 && h <= 8) {   // small rects carry an explicit tile mask (culled for the torch-style weighted sum)
        tm = tile_cull_mask(pr.px, pr.py, pr.sx, pr.sy, vp.k, tx0, ty0, w, h,
                            vp.style == B2S_STYLE_TORCH && vp.exact_bbox == 0);
        my_cnt = __popcll(tm);
      }
      rc = make_uint2((uint32_t)tx0 | ((uint32_t)ty0 << 16), (uint32_t)tx1 | ((uint32_t)ty1 << 16));
    }
    rect[i] = rc;
    tmask[i] = tm;
    dbits[i] = depth_bits(pr.zcam);
    cnt[i] = my_cnt;
    if (rec != nullptr) {
      float4 a, b, c;
      int cm = 0;
      if (pr.ok) {
        float craw[3], dir[3], rinv;
        if (colors != nullptr) {
          float coef[K * 3];
          load_coeffs<K>(colors, i, coef);
          eval_color<K>(vp, coef, mx, my, mz, craw, dir, &rinv);
        } else {
          craw[0] = craw[1] = craw[2] = 0.0f;
        }
        make_record(vp, pr, op, craw, a, b, c, cm);
      } else {
        a = make_float4(0.f, 0.f, (vp.exact_bbox == 0) ? -INFINITY : 0.0f, __int_as_float(0));
        b = make_float4(0.f, 0.f, 0.f, __int_as_float(0));
        c = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      rec[3 * (size_t)i] = a;
      rec[3 * (size_t)i + 1] = b;
      rec[3 * (size_t)i + 2] = c;
      if (cmask_out != nullptr) cmask_out[i] = (uint8_t)cm;
    }
    if (dbg != nullptr) {
      dbg[i] = pr.px; dbg[(size_t)n + i] = pr.py; dbg[2 * (size_t)n + i] = pr.sx;
      dbg[3 * (size_t)n + i] = pr.sy; dbg[4 * (size_t)n + i] = pr.zabs;
    }
    if (dbg_bbox != nullptr) {
      dbg_bbox[4 * (size_t)i] = pr.xmin; dbg_bbox[4 * (size_t)i + 1] = pr.ymin;
      dbg_bbox[4 * (size_t)i + 2] = pr.xmax; dbg_bbox[4 * (size_t)i + 3] = pr.ymax;
    }
  }
  // block sum of tile counts -> bsum[block]
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

int launch_preprocess(const ViewParams& vp, const float* means, const float* scales, const float* colors,
                      const float* opac, int n, float4* rec, uint8_t* cmask, uint2* rect, unsigned long long* tmask,
                      uint32_t* dbits, int* cnt, long long* bsum, float* dbg, int* dbg_bbox, cudaStream_t st) {
  if (n <= 0) return B2S_OK;
  const int blocks = (n + PRE_BLOCK - 1) / PRE_BLOCK;
#define B2S_PRE(KK) preprocess_kernel<KK><<<blocks, PRE_BLOCK, 0, st>>>(vp, means, scales, colors, opac, n, rec, cmask, rect, tmask, dbits, cnt, bsum, dbg, dbg_bbox)
  switch (vp.sh) {
    case 1: B2S_PRE(1); break;
    case 4: B2S_PRE(4); break;
    case 9: B2S_PRE(9); break;
    case 16: B2S_PRE(16); break;
    default: set_error("sh_coeffs must be 1, 4, 9 or 16 (got %d)", vp.sh); return B2S_ERR_INVALID;
  }
#undef B2S_PRE
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

// Forward for ALL local views of a fit iteration in one launch: the parameters -- above all the (N,K,3)
// coefficients, 192 B per Gaussian at K = 16, the dominant HBM read of the per-view kernel -- are read once
// and kept in registers while the view loop writes each view's blend record, clamp mask, tile rect and tile
// mask (the inputs of the counting-sort path; cnt / dbits / bsum belong to the radix path and are not produced).
// HBM: (28 + 12K) B read per Gaussian + 65 B written per Gaussian*view.
constexpr int PRE_VIEWS_SMEM = 32;

template <int K>
__global__ void __launch_bounds__(PRE_BLOCK)
preprocess_views_kernel(const ViewParams* __restrict__ views, int num_views, const float* __restrict__ means,
                        const float* __restrict__ scales, const float* __restrict__ colors,
                        const float* __restrict__ opac, int n, char* __restrict__ prepared, PreparedLayout L) {
  __shared__ ViewParams sv[PRE_VIEWS_SMEM];
  const int i = blockIdx.x * PRE_BLOCK + threadIdx.x;
  const bool live = i < n;
  const int ii = live ? i : 0;
  const float mx = __ldg(means + 3 * (size_t)ii), my = __ldg(means + 3 * (size_t)ii + 1),
              mz = __ldg(means + 3 * (size_t)ii + 2);
  const float raw_s0 = __ldg(scales + 3 * (size_t)ii), raw_s1 = __ldg(scales + 3 * (size_t)ii + 1);
  const float raw_op = __ldg(opac + ii);
  float coef[K * 3];
  load_coeffs<K>(colors, ii, coef);
  // activations are view independent (every view of a fit shares act_flags)
  const ViewParams& v0 = views[0];
  const float s0 = act_scale(v0, raw_s0), s1 = act_scale(v0, raw_s1), op = act_opac(v0, raw_op);
  for (int vbase = 0; vbase < num_views; vbase += PRE_VIEWS_SMEM) {
    const int vcount = min(PRE_VIEWS_SMEM, num_views - vbase);
    __syncthreads();
    {
      const int words = vcount * (int)(sizeof(ViewParams) / 4);
      const int* src = reinterpret_cast<const int*>(views + vbase);
      int* dst = reinterpret_cast<int*>(sv);
      for (int q = threadIdx.x; q < words; q += PRE_BLOCK) dst[q] = src[q];
    }
    __syncthreads();
    if (!live) continue;
    for (int vl = 0; vl < vcount; ++vl) {
      const ViewParams& vp = sv[vl];
      char* base = prepared + (size_t)(vbase + vl) * L.total;
      float4* rec = reinterpret_cast<float4*>(base + L.rec);
      const Proj pr = project_gaussian(vp, mx, my, mz, s0, s1, op);
      uint2 rc = make_uint2(1u, 0u);   // empty tile rect (tx1 < tx0) for culled Gaussians
      unsigned long long tm = 0ull;
      float4 a = make_float4(0.f, 0.f, (vp.exact_bbox == 0) ? -INFINITY : 0.0f, __int_as_float(0)),
             b = make_float4(0.f, 0.f, 0.f, __int_as_float(0)), c = make_float4(0.f, 0.f, 0.f, 0.f);
      int cm = 0;
      if (pr.ok) {
        const int tx0 = pr.xmin / TILE, tx1 = pr.xmax / TILE, ty0 = pr.ymin / TILE, ty1 = pr.ymax / TILE;
        const int w = tx1 - tx0 + 1, h = ty1 - ty0 + 1;
        if (w <= 8 && h <= 8)
          tm = tile_cull_mask(pr.px, pr.py, pr.sx, pr.sy, vp.k, tx0, ty0, w, h,
                              vp.style == B2S_STYLE_TORCH && vp.exact_bbox == 0);
        rc = make_uint2((uint32_t)tx0 | ((uint32_t)ty0 << 16), (uint32_t)tx1 | ((uint32_t)ty1 << 16));
        float craw[3], dir[3], rinv;
        eval_color<K>(vp, coef, mx, my, mz, craw, dir, &rinv);
        make_record(vp, pr, op, craw, a, b, c, cm);
      }
      rec[3 * (size_t)i] = a;
      rec[3 * (size_t)i + 1] = b;
      rec[3 * (size_t)i + 2] = c;
      reinterpret_cast<uint8_t*>(base + L.cmask)[i] = (uint8_t)cm;
      reinterpret_cast<uint2*>(base + L.rect)[i] = rc;
      reinterpret_cast<unsigned long long*>(base + L.tmask)[i] = tm;
    }
  }
}

int launch_preprocess_views(const ViewParams* views_dev, int num_views, int sh, const float* means, const float* scales,
                            const float* colors, const float* opac, int n, char* prepared, cudaStream_t st) {
  if (n <= 0 || num_views <= 0) return B2S_OK;
  const int blocks = (n + PRE_BLOCK - 1) / PRE_BLOCK;
  const PreparedLayout L = prepared_layout(n);
#define B2S_PREV(KK) preprocess_views_kernel<KK><<<blocks, PRE_BLOCK, 0, st>>>(views_dev, num_views, means, scales, colors, opac, n, prepared, L)
  switch (sh) {
    case 1: B2S_PREV(1); break;
    case 4: B2S_PREV(4); break;
    case 9: B2S_PREV(9); break;
    case 16: B2S_PREV(16); break;
    default: set_error("sh_coeffs must be 1, 4, 9 or 16 (got %d)", sh); return B2S_ERR_INVALID;
  }
#undef B2S_PREV
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

// Backward chain rule, for ONE OR MANY views in a single pass over the parameters.
//   gacc[v][12i..] = {dR,dG,dB,dZ, S,Sx,Sxx,Sy, Syy,-,-,-} from the blend backward of view v, where
//   S = sum w*t, Sx = sum w*t*dx, Sxx = sum w*t*dx^2 (same for y).
// Chain rule of SURVEY Appendix A / autograd of torch_renderer.py:57-104,143-150.
//
// The fit loop used to run this once per view and read-modify-write the whole gradient buffer
// (28+12K floats... 220 B at K=16) every time: 64 views x 0.7 GB.  Here the per-view results stay
// compact (48 B per Gaussian per view) and one launch folds all views: each Gaussian's gradients
// are accumulated in registers over the view loop and written once.
// LPG lanes share a Gaussian: lane `sub` owns SH coefficients [sub*KL, sub*KL+KL), so a warp reads
// and writes the (N,K,3) arrays as contiguous 16-byte pieces (coalesced), and the few cross-lane
// sums (raw colour, d colour / d direction) are quad shuffles.
constexpr int BWD_VIEWS_SMEM = 32;   // views staged in shared memory per chunk (a multiple of 4)

template <int K, int LPG>
__global__ void __launch_bounds__(PRE_BLOCK)
preprocess_bwd_kernel(const ViewParams single, const ViewParams* __restrict__ views, int num_views,
                      const float* __restrict__ means, const float* __restrict__ scales,
                      const float* __restrict__ colors, const float* __restrict__ opac, int n, int first, int count,
                      const float4* __restrict__ gacc, float* __restrict__ g_means, float* __restrict__ g_scales,
                      float* __restrict__ g_colors, float* __restrict__ g_opac, int accumulate) {
  constexpr int KL = K / LPG;       // SH coefficients per lane
  constexpr int CL = KL * 3;        // colour floats per lane
  static_assert(K % LPG == 0 && (LPG == 1 || (CL % 4) == 0), "lane split must keep 16-byte pieces");
  __shared__ ViewParams sv[BWD_VIEWS_SMEM];
  // Gaussians [first, first + count) of the n (n is also the per-view stride of gacc): the fit loop folds the views
  // chunk by chunk so that a chunk's gradients can be all-reduced while the next chunk is computed
  const int gidx = blockIdx.x * PRE_BLOCK + threadIdx.x;
  const int i = first + gidx / LPG, sub = gidx % LPG;
  const bool live = gidx / LPG < count;
  const int ii = live ? i : first;

  const float mx = __ldg(means + 3 * (size_t)ii), my = __ldg(means + 3 * (size_t)ii + 1),
              mz = __ldg(means + 3 * (size_t)ii + 2);
  const float raw_s0 = __ldg(scales + 3 * (size_t)ii), raw_s1 = __ldg(scales + 3 * (size_t)ii + 1);
  const float raw_op = __ldg(opac + ii);
  float coef[CL];
  if (colors != nullptr) {
    const float* cp = colors + (size_t)ii * K * 3 + sub * CL;
    if constexpr (CL % 4 == 0) {
#pragma unroll
      for (int q = 0; q < CL / 4; ++q) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(cp) + q);
        coef[4 * q] = v.x; coef[4 * q + 1] = v.y; coef[4 * q + 2] = v.z; coef[4 * q + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int q = 0; q < CL; ++q) coef[q] = __ldg(cp + q);
    }
  } else {
#pragma unroll
    for (int q = 0; q < CL; ++q) coef[q] = 0.0f;
  }

  float gm[3] = {0.f, 0.f, 0.f}, gs0 = 0.f, gs1 = 0.f, gop = 0.f;
  float gcoef[CL];
#pragma unroll
  for (int q = 0; q < CL; ++q) gcoef[q] = 0.0f;

  for (int vbase = 0; vbase < num_views; vbase += BWD_VIEWS_SMEM) {
    const int vcount = min(BWD_VIEWS_SMEM, num_views - vbase);
    if (views != nullptr) {
      __syncthreads();
      const int words = vcount * (int)(sizeof(ViewParams) / 4);
      const int* src = reinterpret_cast<const int*>(views + vbase);
      int* dst = reinterpret_cast<int*>(sv);
      for (int q = threadIdx.x; q < words; q += PRE_BLOCK) dst[q] = src[q];
      __syncthreads();
    }
    for (int vl = 0; vl < vcount; ++vl) {
      const ViewParams& vp = (views != nullptr) ? sv[vl] : single;
      const float4* ga = gacc + ((size_t)(vbase + vl) * n + ii) * 3;
      const float4 g0 = __ldg(ga), g1 = __ldg(ga + 1), g2 = __ldg(ga + 2);
      const float s0 = act_scale(vp, raw_s0), s1 = act_scale(vp, raw_s1), op = act_opac(vp, raw_op);
      const Proj pr = project_gaussian(vp, mx, my, mz, s0, s1, op);
      const bool ok = pr.ok;      // culled in this view: its gacc row is all zeros (no `continue`: the
                                  // colour block below holds warp shuffles that every lane must reach)
      const float dC[3] = {g0.x, g0.y, g0.z};
      float dZ = g2.x;            // row = {dR, dG, dB, Syy | S, Sx, Sxx, Sy | dZ, colour clamp mask, -, -}
      const float S = g1.x, Sx = g1.y, Sxx = g1.z, Sy = g1.w, Syy = g0.w;

      if (ok) {
        // opacity: w = op * E  =>  dL/dop = S / op
        float go = (op > 0.0f) ? S / op : S;   // at op == 0 the blend backward accumulated sum E*t directly
        if (vp.act & B2S_ACT_OPACITY_SIGMOID) go *= op * (1.0f - op);
        gop += go;

        // position / sigma
        const float isx2 = 1.0f / (pr.sx * pr.sx), isy2 = 1.0f / (pr.sy * pr.sy);
        const float dpx = Sx * isx2, dpy = Sy * isy2;
        const float dsx = Sxx * isx2 / pr.sx, dsy = Syy * isy2 / pr.sy;
        if (pr.ax >= 1.0f) {   // clamp_min(1) passes the gradient on [1, inf)
          const float sgn = (vp.style == B2S_STYLE_TORCH) ? ((s0 > 0.f) - (s0 < 0.f)) : 1.0f;
          float t = dsx * sgn * (0.5f * vp.wf * vp.fx / pr.zabs);
          if (vp.act & B2S_ACT_SCALES_SOFTPLUS) t *= sigmoidf_acc(raw_s0);
          gs0 += t;
          dZ -= dsx * pr.ax / pr.zabs;
        }
        if (pr.ay >= 1.0f) {
          const float sgn = (vp.style == B2S_STYLE_TORCH) ? ((s1 > 0.f) - (s1 < 0.f)) : 1.0f;
          float t = dsy * sgn * (0.5f * vp.hf * vp.fy / pr.zabs);
          if (vp.act & B2S_ACT_SCALES_SOFTPLUS) t *= sigmoidf_acc(raw_s1);
          gs1 += t;
          dZ -= dsy * pr.ay / pr.zabs;
        }
        // zabs = max(|cam.z|, 1e-6)
        float dcam[4] = {0.f, 0.f, 0.f, 0.f};
        if (fabsf(pr.zcam) >= 1e-6f) dcam[2] = dZ * ((pr.zcam > 0.f) ? 1.0f : -1.0f);
        // px,py -> ndc -> clip
        const float dnx = dpx * 0.5f * vp.wm1, dny = -dpy * 0.5f * vp.hm1;
        float dclip[4];
        dclip[0] = dnx / pr.wsafe;
        dclip[1] = dny / pr.wsafe;
        dclip[2] = 0.0f;
        dclip[3] = (fabsf(pr.w) < 1e-8f) ? 0.0f : -(dnx * pr.ndcx + dny * pr.ndcy) / pr.wsafe;
        // clip = P cam ; cam = V [m,1]
  #pragma unroll
        for (int c = 0; c < 4; ++c)
  #pragma unroll
          for (int r = 0; r < 4; ++r) dcam[c] = fmaf(vp.proj[4 * r + c], dclip[r], dcam[c]);
  #pragma unroll
        for (int c = 0; c < 3; ++c)
  #pragma unroll
          for (int r = 0; r < 4; ++r) gm[c] = fmaf(vp.view[4 * r + c], dcam[r], gm[c]);

      }
      // colour
      if (colors != nullptr) {
        if constexpr (K == 1) {
          if (!ok) continue;
          float craw[3] = {coef[0], coef[1], coef[2]};
          if (vp.act & B2S_ACT_COLORS_SIGMOID) {
#pragma unroll
            for (int q = 0; q < 3; ++q) craw[q] = sigmoidf_acc(craw[q]);
          }
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            const float dc = (craw[q] >= 0.0f && craw[q] <= 1.0f) ? dC[q] : 0.0f;
            gcoef[q] += (vp.act & B2S_ACT_COLORS_SIGMOID) ? dc * craw[q] * (1.0f - craw[q]) : dc;
          }
        } else {
          const float vx = vp.cam[0] - mx, vy = vp.cam[1] - my, vz = vp.cam[2] - mz;
          const float r = sqrtf(vx * vx + vy * vy + vz * vz);
          const float rinv = 1.0f / (r + 1e-8f);
          const float dx = vx * rinv, dy = vy * rinv, dz = vz * rinv;
          float b[16], bx[16], by[16], bz[16];
          sh_basis(dx, dy, dz, K, b);
          sh_basis_grad(dx, dy, dz, K, bx, by, bz);
          // this lane's KL basis functions (static register indexing: select, do not index by `sub`)
          float lb[KL], lbx[KL], lby[KL], lbz[KL];
#pragma unroll
          for (int j = 0; j < KL; ++j) {
            lb[j] = b[j]; lbx[j] = bx[j]; lby[j] = by[j]; lbz[j] = bz[j];
#pragma unroll
            for (int s2 = 1; s2 < LPG; ++s2) {
              if (sub == s2) { lb[j] = b[s2 * KL + j]; lbx[j] = bx[s2 * KL + j]; lby[j] = by[s2 * KL + j]; lbz[j] = bz[s2 * KL + j]; }
            }
          }
          // colour clamp mask (torch_renderer.py:144) planted next to the sums by gacc_init_kernel
          const int cmask = __float_as_int(g2.y);
          const float dc0 = (ok && (cmask & 1)) ? dC[0] : 0.0f, dc1 = (ok && (cmask & 2)) ? dC[1] : 0.0f,
                      dc2 = (ok && (cmask & 4)) ? dC[2] : 0.0f;
          float dd0 = 0.f, dd1 = 0.f, dd2 = 0.f;
#pragma unroll
          for (int j = 0; j < KL; ++j) {
            const float s = coef[3 * j] * dc0 + coef[3 * j + 1] * dc1 + coef[3 * j + 2] * dc2;
            gcoef[3 * j] = fmaf(lb[j], dc0, gcoef[3 * j]);
            gcoef[3 * j + 1] = fmaf(lb[j], dc1, gcoef[3 * j + 1]);
            gcoef[3 * j + 2] = fmaf(lb[j], dc2, gcoef[3 * j + 2]);
            dd0 = fmaf(lbx[j], s, dd0);
            dd1 = fmaf(lby[j], s, dd1);
            dd2 = fmaf(lbz[j], s, dd2);
          }
#pragma unroll
          for (int o = 1; o < LPG; o <<= 1) {
            dd0 += __shfl_xor_sync(0xffffffffu, dd0, o);
            dd1 += __shfl_xor_sync(0xffffffffu, dd1, o);
            dd2 += __shfl_xor_sync(0xffffffffu, dd2, o);
          }
          // d = v/(r+eps), v = cam - m :  dv = dd/(r+eps) - v (v.dd) / (r (r+eps)^2) ; dm = -dv
          const float vdd = vx * dd0 + vy * dd1 + vz * dd2;
          const float k2 = (r > 0.0f) ? vdd * rinv * rinv / r : 0.0f;
          if (ok) {
            gm[0] -= dd0 * rinv - vx * k2;
            gm[1] -= dd1 * rinv - vy * k2;
            gm[2] -= dd2 * rinv - vz * k2;
          }
        }
      }
    }
  }
  if (!live) return;
  if (sub == 0) {
    if (accumulate) {
      g_means[3 * (size_t)i] += gm[0]; g_means[3 * (size_t)i + 1] += gm[1]; g_means[3 * (size_t)i + 2] += gm[2];
      g_scales[3 * (size_t)i] += gs0; g_scales[3 * (size_t)i + 1] += gs1;
      g_opac[i] += gop;
    } else {
      g_means[3 * (size_t)i] = gm[0]; g_means[3 * (size_t)i + 1] = gm[1]; g_means[3 * (size_t)i + 2] = gm[2];
      g_scales[3 * (size_t)i] = gs0; g_scales[3 * (size_t)i + 1] = gs1; g_scales[3 * (size_t)i + 2] = 0.0f;
      g_opac[i] = gop;
    }
  }
  if (g_colors != nullptr) {
    float* gp = g_colors + (size_t)i * K * 3 + sub * CL;
    if constexpr (CL % 4 == 0) {
#pragma unroll
      for (int q = 0; q < CL / 4; ++q) {
        float4 v = make_float4(gcoef[4 * q], gcoef[4 * q + 1], gcoef[4 * q + 2], gcoef[4 * q + 3]);
        if (accumulate) {
          const float4 o = reinterpret_cast<const float4*>(gp)[q];
          v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
        }
        reinterpret_cast<float4*>(gp)[q] = v;
      }
    } else {
#pragma unroll
      for (int q = 0; q < CL; ++q) gp[q] = accumulate ? gp[q] + gcoef[q] : gcoef[q];
    }
  }
}

// ---- K = 16 (SH degree 3): warp-uniform coefficient groups ----------------------------------------
// The generic kernel above lets 4 lanes share a Gaussian, which makes every lane repeat the projection chain and
// evaluate all 16 basis functions + 48 derivatives to use 4 of them.  Here the split is per WARP: a block of 8
// warps covers 64 Gaussians x 4 coefficient groups, warp w handles group (w & 3) of Gaussians (w >> 2)*32 + lane.
// `sub` is warp uniform, so each warp evaluates only its own 4 basis functions (compile-time indices, dead code
// eliminated).  The projection / sigma / opacity / position chain of view v runs on the warp with (v & 3) == sub,
// so the four warps of a Gaussian carry the same load.  Every gradient is linear in the per-warp partial sums:
// each warp keeps private d/dmean, d/dscale, d/dopacity sums and the four are added through shared memory once,
// after the view loop.
template <int SUB>
__device__ __forceinline__ void sh16_group_accumulate(const ViewParams& vp, float mx, float my, float mz,
                                                      const float (&coef)[12], float dc0, float dc1, float dc2,
                                                      float (&gcoef)[12], float (&gm)[3]) {
  const float vx = vp.cam[0] - mx, vy = vp.cam[1] - my, vz = vp.cam[2] - mz;
  const float r = sqrtf(vx * vx + vy * vy + vz * vz);
  const float rinv = __frcp_rn(r + 1e-8f);
  const float dx = vx * rinv, dy = vy * rinv, dz = vz * rinv;
  float b[16], bx[16], by[16], bz[16];
  sh_basis(dx, dy, dz, 16, b);
  sh_basis_grad(dx, dy, dz, 16, bx, by, bz);
  float dd0 = 0.f, dd1 = 0.f, dd2 = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    constexpr int K0 = 4 * SUB;
    const float s = coef[3 * j] * dc0 + coef[3 * j + 1] * dc1 + coef[3 * j + 2] * dc2;
    gcoef[3 * j] = fmaf(b[K0 + j], dc0, gcoef[3 * j]);
    gcoef[3 * j + 1] = fmaf(b[K0 + j], dc1, gcoef[3 * j + 1]);
    gcoef[3 * j + 2] = fmaf(b[K0 + j], dc2, gcoef[3 * j + 2]);
    dd0 = fmaf(bx[K0 + j], s, dd0);
    dd1 = fmaf(by[K0 + j], s, dd1);
    dd2 = fmaf(bz[K0 + j], s, dd2);
  }
  // d = v/(r+eps), v = cam - m :  dv = dd/(r+eps) - v (v.dd) / (r (r+eps)^2) ; dm = -dv
  const float vdd = vx * dd0 + vy * dd1 + vz * dd2;
  const float k2 = (r > 0.0f) ? vdd * rinv * rinv * __frcp_rn(r) : 0.0f;
  gm[0] -= dd0 * rinv - vx * k2;
  gm[1] -= dd1 * rinv - vy * k2;
  gm[2] -= dd2 * rinv - vz * k2;
}

__global__ void __launch_bounds__(PRE_BLOCK)
preprocess_bwd_sh16_kernel(const ViewParams single, const ViewParams* __restrict__ views, int num_views,
                           const float* __restrict__ means, const float* __restrict__ scales,
                           const float* __restrict__ colors, const float* __restrict__ opac, int n, int first, int count,
                           const float4* __restrict__ gacc, float* __restrict__ g_means, float* __restrict__ g_scales,
                           float* __restrict__ g_colors, float* __restrict__ g_opac, int accumulate) {
  static_assert(PRE_BLOCK == 256, "8 warps: 2 x 32 Gaussians x 4 coefficient groups");
  __shared__ ViewParams sv[BWD_VIEWS_SMEM];
  __shared__ float part[3][6][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = warp & 3, slot = (warp >> 2) * 32 + lane;
  const int i = first + blockIdx.x * 64 + slot;
  const bool live = (int)blockIdx.x * 64 + slot < count;
  const int ii = live ? i : first;

  const float mx = __ldg(means + 3 * (size_t)ii), my = __ldg(means + 3 * (size_t)ii + 1),
              mz = __ldg(means + 3 * (size_t)ii + 2);
  // activations are view independent (every view of a fit shares act_flags): once per Gaussian, not per view
  const ViewParams& v0 = (views != nullptr) ? views[0] : single;
  const int act = v0.act;
  float s0, s1, op, ds0 = 1.f, ds1 = 1.f, dop = 1.f;   // activated values, d activation / d raw
  {
    const float raw_s0 = __ldg(scales + 3 * (size_t)ii), raw_s1 = __ldg(scales + 3 * (size_t)ii + 1);
    const float raw_op = __ldg(opac + ii);
    s0 = raw_s0; s1 = raw_s1; op = raw_op;
    if (act & B2S_ACT_SCALES_SOFTPLUS) {
      s0 = softplusf_acc(raw_s0) + 1e-3f; s1 = softplusf_acc(raw_s1) + 1e-3f;
      ds0 = sigmoidf_acc(raw_s0); ds1 = sigmoidf_acc(raw_s1);
    }
    if (act & B2S_ACT_OPACITY_SIGMOID) { op = sigmoidf_acc(raw_op); dop = op * (1.0f - op); }
  }
  const float inv_op = (op > 0.0f) ? dop / op : dop;   // at op == 0 the blend backward accumulated sum E*t directly
  float coef[12], gcoef[12];
  {
    const float4* cp = reinterpret_cast<const float4*>(colors + (size_t)ii * 48 + sub * 12);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const float4 v = __ldg(cp + q);
      coef[4 * q] = v.x; coef[4 * q + 1] = v.y; coef[4 * q + 2] = v.z; coef[4 * q + 3] = v.w;
    }
#pragma unroll
    for (int q = 0; q < 12; ++q) gcoef[q] = 0.0f;
  }
  float gm[3] = {0.f, 0.f, 0.f}, gs0 = 0.f, gs1 = 0.f, gop = 0.f;

  for (int vbase = 0; vbase < num_views; vbase += BWD_VIEWS_SMEM) {
    const int vcount = min(BWD_VIEWS_SMEM, num_views - vbase);
    if (views != nullptr) {
      __syncthreads();
      const int words = vcount * (int)(sizeof(ViewParams) / 4);
      const int* src = reinterpret_cast<const int*>(views + vbase);
      int* dst = reinterpret_cast<int*>(sv);
      for (int q = threadIdx.x; q < words; q += PRE_BLOCK) dst[q] = src[q];
      __syncthreads();
    }
    // the next view's sums are fetched while this view's chain runs (the loop is otherwise one long
    // dependent chain behind three 16-byte loads)
    const float4* ga = gacc + ((size_t)vbase * n + ii) * 3;
    float4 g0 = __ldg(ga), g1 = make_float4(0.f, 0.f, 0.f, 0.f), g2 = __ldg(ga + 2);
    if (sub == 0) g1 = __ldg(ga + 1);                  // BWD_VIEWS_SMEM % 4 == 0: view vbase belongs to group 0
    for (int vl = 0; vl < vcount; ++vl) {
      const ViewParams& vp = (views != nullptr) ? sv[vl] : single;
      const float4 c0 = g0, c1 = g1, c2 = g2;
      const bool mine = (vl & 3) == sub;               // warp uniform: this warp runs the projection chain of the view
      if (vl + 1 < vcount) {
        const float4* gn = gacc + ((size_t)(vbase + vl + 1) * n + ii) * 3;
        g0 = __ldg(gn); g2 = __ldg(gn + 2);
        if (((vl + 1) & 3) == sub) g1 = __ldg(gn + 1);
      }
      // c0 = {dR, dG, dB, Syy} (zero row when the Gaussian is culled in this view), c2 = {dZ, colour clamp mask, -, -}
      const int cmask = __float_as_int(c2.y);
      const float dc0 = (cmask & 1) ? c0.x : 0.0f, dc1 = (cmask & 2) ? c0.y : 0.0f, dc2 = (cmask & 4) ? c0.z : 0.0f;
      if (mine) {                      // projection -> sigma / opacity / position chain
        const Proj pr = project_gaussian(vp, mx, my, mz, s0, s1, op);
        if (pr.ok) {
          float dZ = c2.x;
          const float S = c1.x, Sx = c1.y, Sxx = c1.z, Sy = c1.w, Syy = c0.w;
          gop = fmaf(S, inv_op, gop);
          const float rsx = __frcp_rn(pr.sx), rsy = __frcp_rn(pr.sy), rz = __frcp_rn(pr.zabs);
          const float isx2 = rsx * rsx, isy2 = rsy * rsy;
          const float dpx = Sx * isx2, dpy = Sy * isy2;
          const float dsx = Sxx * isx2 * rsx, dsy = Syy * isy2 * rsy;
          if (pr.ax >= 1.0f) {         // clamp_min(1) passes the gradient on [1, inf)
            const float sgn = (vp.style == B2S_STYLE_TORCH) ? ((s0 > 0.f) - (s0 < 0.f)) : 1.0f;
            gs0 = fmaf(dsx * sgn, 0.5f * vp.wf * vp.fx * rz, gs0);
            dZ -= dsx * pr.ax * rz;
          }
          if (pr.ay >= 1.0f) {
            const float sgn = (vp.style == B2S_STYLE_TORCH) ? ((s1 > 0.f) - (s1 < 0.f)) : 1.0f;
            gs1 = fmaf(dsy * sgn, 0.5f * vp.hf * vp.fy * rz, gs1);
            dZ -= dsy * pr.ay * rz;
          }
          float dcam[4] = {0.f, 0.f, 0.f, 0.f};
          if (fabsf(pr.zcam) >= 1e-6f) dcam[2] = dZ * ((pr.zcam > 0.f) ? 1.0f : -1.0f);
          const float rw = __frcp_rn(pr.wsafe);
          const float dnx = dpx * 0.5f * vp.wm1, dny = -dpy * 0.5f * vp.hm1;
          float dclip[4];
          dclip[0] = dnx * rw;
          dclip[1] = dny * rw;
          dclip[2] = 0.0f;
          dclip[3] = (fabsf(pr.w) < 1e-8f) ? 0.0f : -(dnx * pr.ndcx + dny * pr.ndcy) * rw;
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int r = 0; r < 4; ++r) dcam[c] = fmaf(vp.proj[4 * r + c], dclip[r], dcam[c]);
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int r = 0; r < 4; ++r) gm[c] = fmaf(vp.view[4 * r + c], dcam[r], gm[c]);
        }
      }
      // colour: dC is zero for culled Gaussians and masked by the clamp above
      switch (sub) {
        case 0: sh16_group_accumulate<0>(vp, mx, my, mz, coef, dc0, dc1, dc2, gcoef, gm); break;
        case 1: sh16_group_accumulate<1>(vp, mx, my, mz, coef, dc0, dc1, dc2, gcoef, gm); break;
        case 2: sh16_group_accumulate<2>(vp, mx, my, mz, coef, dc0, dc1, dc2, gcoef, gm); break;
        default: sh16_group_accumulate<3>(vp, mx, my, mz, coef, dc0, dc1, dc2, gcoef, gm); break;
      }
    }
  }
  gs0 *= ds0;    // d softplus / d raw, applied once to the view sums
  gs1 *= ds1;
  // fold the four warps' partial sums
  if (sub != 0) {
    float* dst = &part[sub - 1][0][slot];
    dst[0] = gm[0]; dst[64] = gm[1]; dst[128] = gm[2]; dst[192] = gs0; dst[256] = gs1; dst[320] = gop;
  }
  __syncthreads();
  if (!live) return;
  if (sub == 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) gm[c] += part[0][c][slot] + part[1][c][slot] + part[2][c][slot];
    gs0 += part[0][3][slot] + part[1][3][slot] + part[2][3][slot];
    gs1 += part[0][4][slot] + part[1][4][slot] + part[2][4][slot];
    gop += part[0][5][slot] + part[1][5][slot] + part[2][5][slot];
    if (accumulate) {
      g_means[3 * (size_t)i] += gm[0]; g_means[3 * (size_t)i + 1] += gm[1]; g_means[3 * (size_t)i + 2] += gm[2];
      g_scales[3 * (size_t)i] += gs0; g_scales[3 * (size_t)i + 1] += gs1;
      g_opac[i] += gop;
    } else {
      g_means[3 * (size_t)i] = gm[0]; g_means[3 * (size_t)i + 1] = gm[1]; g_means[3 * (size_t)i + 2] = gm[2];
      g_scales[3 * (size_t)i] = gs0; g_scales[3 * (size_t)i + 1] = gs1; g_scales[3 * (size_t)i + 2] = 0.0f;
      g_opac[i] = gop;
    }
  }
  if (g_colors != nullptr) {
    float4* gp = reinterpret_cast<float4*>(g_colors + (size_t)i * 48 + sub * 12);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      float4 v = make_float4(gcoef[4 * q], gcoef[4 * q + 1], gcoef[4 * q + 2], gcoef[4 * q + 3]);
      if (accumulate) {
        const float4 o = gp[q];
        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
      }
      gp[q] = v;
    }
  }
}

int launch_preprocess_bwd(const ViewParams* single, const ViewParams* views_dev, int num_views, int sh,
                          const float* means, const float* scales, const float* colors, const float* opac, int n,
                          int first, int count, const float* gacc, float* g_means, float* g_scales, float* g_colors,
                          float* g_opac, int accumulate, cudaStream_t st) {
  if (n <= 0 || num_views <= 0 || count <= 0) return B2S_OK;
  ViewParams one;
  if (single != nullptr) one = *single; else memset(&one, 0, sizeof(one));
#define B2S_PREB(KK, LL)                                                                                              \
  preprocess_bwd_kernel<KK, LL><<<(int)(((size_t)count * LL + PRE_BLOCK - 1) / PRE_BLOCK), PRE_BLOCK, 0, st>>>(       \
      one, views_dev, num_views, means, scales, colors, opac, n, first, count, reinterpret_cast<const float4*>(gacc),  \
      g_means, g_scales, g_colors, g_opac, accumulate)
  switch (sh) {
    case 1: B2S_PREB(1, 1); break;
    case 4: B2S_PREB(4, 1); break;
    case 9: B2S_PREB(9, 1); break;
    case 16:
      if (colors != nullptr && g_colors != nullptr) {
        preprocess_bwd_sh16_kernel<<<(count + 63) / 64, PRE_BLOCK, 0, st>>>(one, views_dev, num_views, means, scales, colors, opac, n,
                                                                       first, count, reinterpret_cast<const float4*>(gacc), g_means, g_scales,
                                                                       g_colors, g_opac, accumulate);
      } else {
        B2S_PREB(16, 4);
      }
      break;
    default: set_error("sh_coeffs must be 1, 4, 9 or 16 (got %d)", sh); return B2S_ERR_INVALID;
  }
#undef B2S_PREB
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

}  // namespace b2s
