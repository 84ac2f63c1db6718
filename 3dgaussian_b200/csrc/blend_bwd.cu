// Backward of the weighted-sum blend.  The reference has no hand-written backward (torch
// autograd over python/torch_renderer.py:143-203, O(N*H*W) saved activations); this is the
// closed form of SURVEY Appendix A.
//
// Because the weighted sum has no transmittance, dL/dw_i(p) = gA(p).c_i + gW(p) + gD(p) z_i
// depends on five per-pixel numbers only.  Schedule: one CTA per tile; the tile's per-pixel
// g-buffer (256 x 5 floats) is computed once into shared memory from the saved accumulators
// and the incoming image gradients; then ONE THREAD OWNS ONE GAUSSIAN of the tile's list and
// sweeps the 256 pixels, reading g by shared-memory broadcast and keeping its nine partial
// sums in registers.  There is no cross-thread reduction at all (a pixel-per-thread schedule
// needs ~45 shuffles per Gaussian per warp); each thread finishes with three 16-byte vector
// reductions (red.global.add.v4.f32) into the per-Gaussian accumulator.
//
// Per pixel-pair: 1 FADD + 1 MUFU.EX2 + 10 FP32 (+3 with depth gradients) + 1 LDS.128.
#include "common.cuh"

namespace b2s {

constexpr int BB_THREADS = 128;

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

template <bool DEPTH>
__global__ void __launch_bounds__(BB_THREADS)
blend_wsum_bwd_kernel(const ViewParams vp, const float4* __restrict__ rec, const int* __restrict__ vals,
                      const int2* __restrict__ ranges, const float* __restrict__ acc,
                      const float* __restrict__ g_rgb, const float* __restrict__ g_alpha,
                      const float* __restrict__ g_depth, float* __restrict__ gacc) {
  __shared__ __align__(16) float4 sG[TILE_PIX];   // gA.r, gA.g, gA.b, gW
  __shared__ __align__(16) float sGD[TILE_PIX];   // gD
  const int tile = blockIdx.x;
  const int2 rg = ranges[tile];
  const int n = rg.y - rg.x;
  if (n <= 0) return;
  const int tx = tile % vp.tiles_x, ty = tile / vp.tiles_x;
  const size_t hw = (size_t)vp.width * vp.height;

  // ---- per-pixel g-buffer ("per pixel") ----
  for (int q = threadIdx.x; q < TILE_PIX; q += BB_THREADS) {
    const int xi = tx * TILE + (q & 15), yi = ty * TILE + (q >> 4);
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    float gd = 0.f;
    if (xi < vp.width && yi < vp.height) {
      const size_t p = (size_t)yi * vp.width + xi;
      const float A0 = acc[p], A1 = acc[hw + p], A2 = acc[2 * hw + p], W = acc[3 * hw + p];
      const float inv = 1.0f / (1.0f + W);
      const float o0 = (vp.bg[0] + A0) * inv, o1 = (vp.bg[1] + A1) * inv, o2 = (vp.bg[2] + A2) * inv;
      g.x = (o0 >= 0.f && o0 <= 1.f) ? g_rgb[3 * p] * inv : 0.f;
      g.y = (o1 >= 0.f && o1 <= 1.f) ? g_rgb[3 * p + 1] * inv : 0.f;
      g.z = (o2 >= 0.f && o2 <= 1.f) ? g_rgb[3 * p + 2] * inv : 0.f;
      g.w = -(g.x * o0 + g.y * o1 + g.z * o2);
      if (g_alpha != nullptr) g.w = fmaf(g_alpha[p], inv * inv, g.w);
      if (DEPTH) {
        const float D = acc[4 * hw + p];
        const float iw = 1.0f / (W + 1e-6f);
        const float gdep = (D * iw >= 0.f) ? g_depth[p] : 0.f;
        gd = gdep * iw;
        g.w = fmaf(-gdep * D, iw * iw, g.w);
      }
    }
    sG[q] = g;
    sGD[q] = gd;
  }
  __syncthreads();

  const float x0 = tx * TILE + 0.5f, y0 = ty * TILE + 0.5f;
  for (int base = 0; base < n; base += BB_THREADS) {
    const int i = base + threadIdx.x;
    const bool active = i < n;
    int id = 0;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = make_float4(0.f, 0.f, 0.f, -INFINITY);
    float z = 0.f;
    if (active) {
      id = __ldg(vals + rg.x + i);
      a = __ldg(rec + 3 * (size_t)id);
      b = __ldg(rec + 3 * (size_t)id + 1);
      z = __ldg(reinterpret_cast<const float*>(rec + 3 * (size_t)id + 2));
    }
    // op == 0 (log2 op = -inf): the forward weight is 0 but torch's clamp_min(0) still passes
    // dL/dop = sum E*t at 0, so sweep with E instead of w and keep only S.
    const bool zero_op = (b.w == -INFINITY);
    const float lop = zero_op ? 0.0f : b.w;
    const float dx0 = x0 - a.x;
    float ex[TILE], colS[TILE];
#pragma unroll
    for (int c = 0; c < TILE; ++c) {
      const float dx = dx0 + (float)c;
      ex[c] = fmaf(a.z * dx, dx, lop);   // qx dx^2 + log2(op)
      colS[c] = 0.f;
    }
    float dR = 0.f, dG = 0.f, dB = 0.f, dZ = 0.f, S = 0.f, Sy = 0.f, Syy = 0.f;
#pragma unroll 1
    for (int r = 0; r < TILE; ++r) {
      const float dy = (y0 + (float)r) - a.y;
      const float ey = a.w * dy * dy;
      float rowS = 0.f;
      float gdv[TILE];
      if (DEPTH) {
#pragma unroll
        for (int c4 = 0; c4 < TILE / 4; ++c4) {
          const float4 t4 = reinterpret_cast<const float4*>(sGD)[r * (TILE / 4) + c4];
          gdv[4 * c4] = t4.x; gdv[4 * c4 + 1] = t4.y; gdv[4 * c4 + 2] = t4.z; gdv[4 * c4 + 3] = t4.w;
        }
      }
#pragma unroll
      for (int c = 0; c < TILE; ++c) {
        const float w = ex2_approx(ex[c] + ey);
        const float4 g = sG[r * TILE + c];
        float t = fmaf(g.x, b.x, fmaf(g.y, b.y, fmaf(g.z, b.z, g.w)));
        if (DEPTH) {
          t = fmaf(gdv[c], z, t);
          dZ = fmaf(w, gdv[c], dZ);
        }
        const float de = w * t;
        dR = fmaf(w, g.x, dR);
        dG = fmaf(w, g.y, dG);
        dB = fmaf(w, g.z, dB);
        rowS += de;
        colS[c] += de;
      }
      S += rowS;
      const float rd = rowS * dy;
      Sy += rd;
      Syy = fmaf(rd, dy, Syy);
    }
    float Sx = 0.f, Sxx = 0.f;
#pragma unroll
    for (int c = 0; c < TILE; ++c) {
      const float dx = dx0 + (float)c;
      const float cd = colS[c] * dx;
      Sx += cd;
      Sxx = fmaf(cd, dx, Sxx);
    }
    if (zero_op) dR = dG = dB = dZ = Sx = Sxx = Sy = Syy = 0.0f;
    if (active) {
      float* dst = gacc + (size_t)id * GACC_F;
      red_add_v4(dst, dR, dG, dB, dZ);
      red_add_v4(dst + 4, S, Sx, Sxx, Sy);
      atomicAdd(dst + 8, Syy);
    }
  }
}

int launch_blend_wsum_bwd(const ViewParams& vp, const float4* rec, const int* vals, const int2* ranges,
                          const float* acc, const float* g_rgb, const float* g_alpha, const float* g_depth,
                          float* gacc, cudaStream_t st) {
  if (vp.n_tiles <= 0) return B2S_OK;
  if (g_depth != nullptr)
    blend_wsum_bwd_kernel<true><<<vp.n_tiles, BB_THREADS, 0, st>>>(vp, rec, vals, ranges, acc, g_rgb, g_alpha, g_depth, gacc);
  else
    blend_wsum_bwd_kernel<false><<<vp.n_tiles, BB_THREADS, 0, st>>>(vp, rec, vals, ranges, acc, g_rgb, g_alpha, g_depth, gacc);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

}  // namespace b2s
