// Backward of the weighted-sum blend.  The reference has no hand-written backward (torch
// autograd over python/torch_renderer.py:143-203, O(N*H*W) saved activations); this is the
// closed form of SURVEY Appendix A.
//
// Because the weighted sum has no transmittance, dL/dw_i(p) = gA(p).c_i + gW(p) + gD(p) z_i
// depends on five per-pixel numbers only.  Schedule:
//   1. gbuf_kernel: per-pixel g-buffer (gA.rgb, gW, gD) from the saved accumulators and the
//      incoming image gradients, stored tile-major (4 KB + 1 KB contiguous per tile).
//   2. blend_wsum_bwd_kernel: one CTA per work unit (tile, <= SEG Gaussians).  The tile's g-buffer
//      is fetched with ONE TMA bulk copy (cp.async.bulk + mbarrier) into shared memory; then ONE
//      THREAD OWNS ONE GAUSSIAN and sweeps the 256 pixels, reading g by shared-memory broadcast
//      and keeping its nine partial sums in registers.  No cross-thread reduction (a
//      pixel-per-thread schedule needs ~45 shuffles per Gaussian per warp); each thread ends with
//      vector reductions (red.global.add.v4.f32) into the 48-byte per-Gaussian accumulator.
//
// Per pixel-pair: 1 FADD + 1 MUFU.EX2 + 10 FP32 (+2 with depth gradients) + 1 LDS.128.
#include "common.cuh"

namespace b2s {

constexpr int BB_THREADS = 128;

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// ---- per-pixel g-buffer ("per pixel") --------------------------------------
template <bool DEPTH>
__global__ void __launch_bounds__(TILE_PIX)
gbuf_kernel(const ViewParams vp, const float* __restrict__ acc, const float* __restrict__ g_rgb,
            const float* __restrict__ g_alpha, const float* __restrict__ g_depth, float4* __restrict__ gbuf4,
            float* __restrict__ gbufd) {
  const int tile = blockIdx.x, q = threadIdx.x;
  const int xi = (tile % vp.tiles_x) * TILE + (q & 15), yi = (tile / vp.tiles_x) * TILE + (q >> 4);
  const size_t hw = (size_t)vp.width * vp.height;
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
  gbuf4[(size_t)tile * TILE_PIX + q] = g;
  if (DEPTH) gbufd[(size_t)tile * TILE_PIX + q] = gd;
}

// ---- TMA bulk copy helpers (global -> shared, completion on an mbarrier) ------------------------
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(d),
               "l"(gmem_src), "r"(bytes), "r"(b)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  unsigned done = 0;
  for (int spin = 0; spin < (1 << 20); ++spin) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();   // a bulk copy that never lands is a bug: fail the launch instead of hanging the GPU
}

template <bool DEPTH>
__global__ void __launch_bounds__(BB_THREADS)
blend_wsum_bwd_kernel(const ViewParams vp, const float4* __restrict__ rec, const int* __restrict__ vals,
                      const int2* __restrict__ ranges, const int* __restrict__ unit_start,
                      const int2* __restrict__ units, const float4* __restrict__ gbuf4,
                      const float* __restrict__ gbufd, float* __restrict__ gacc) {
  __shared__ __align__(128) float4 sG[TILE_PIX];   // gA.r, gA.g, gA.b, gW
  __shared__ __align__(128) float sGD[TILE_PIX];   // gD
  __shared__ __align__(8) unsigned long long bar;
  const int u = blockIdx.x;
  if (u >= unit_start[vp.n_tiles]) return;
  const int2 ud = units[u];
  const int tile = ud.x;
  const int2 rg = ranges[tile];
  const int start = rg.x + ud.y * SEG;
  const int n = min(SEG, rg.y - start);
  if (n <= 0) return;
  const int tx = tile % vp.tiles_x, ty = tile / vp.tiles_x;

  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, TILE_PIX * 16 + (DEPTH ? TILE_PIX * 4 : 0));
    bulk_g2s(sG, gbuf4 + (size_t)tile * TILE_PIX, TILE_PIX * 16, &bar);
    if (DEPTH) bulk_g2s(sGD, gbufd + (size_t)tile * TILE_PIX, TILE_PIX * 4, &bar);
  }
  // overlap the record gather with the bulk copy
  const float x0 = tx * TILE + 0.5f, y0 = ty * TILE + 0.5f;
  bool waited = false;
  for (int base = 0; base < n; base += BB_THREADS) {
    const int i = base + threadIdx.x;
    const bool active = i < n;
    int id = 0;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = make_float4(0.f, 0.f, 0.f, -INFINITY);
    float z = 0.f;
    if (active) {
      id = __ldg(vals + start + i);
      a = __ldg(rec + 3 * (size_t)id);
      b = __ldg(rec + 3 * (size_t)id + 1);
      if (DEPTH) z = __ldg(reinterpret_cast<const float*>(rec + 3 * (size_t)id + 2));
    }
    if (!waited) {
      mbar_wait(&bar, 0);
      waited = true;
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
                          const int* unit_start, const int2* units, int64_t unit_cap, const float* acc,
                          const float* g_rgb, const float* g_alpha, const float* g_depth, float* gbuf,
                          float* gacc, cudaStream_t st) {
  if (vp.n_tiles <= 0) return B2S_OK;
  float4* gbuf4 = reinterpret_cast<float4*>(gbuf);
  float* gbufd = gbuf + (size_t)vp.n_tiles * TILE_PIX * 4;
  if (g_depth != nullptr) {
    gbuf_kernel<true><<<vp.n_tiles, TILE_PIX, 0, st>>>(vp, acc, g_rgb, g_alpha, g_depth, gbuf4, gbufd);
    B2S_LAUNCH_CHECK();
    blend_wsum_bwd_kernel<true><<<(int)unit_cap, BB_THREADS, 0, st>>>(vp, rec, vals, ranges, unit_start, units, gbuf4, gbufd, gacc);
  } else {
    gbuf_kernel<false><<<vp.n_tiles, TILE_PIX, 0, st>>>(vp, acc, g_rgb, g_alpha, g_depth, gbuf4, gbufd);
    B2S_LAUNCH_CHECK();
    blend_wsum_bwd_kernel<false><<<(int)unit_cap, BB_THREADS, 0, st>>>(vp, rec, vals, ranges, unit_start, units, gbuf4, gbufd, gacc);
  }
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

}  // namespace b2s
