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
// The weight is separable (axis-aligned Gaussians): w(r,c) = fx[c] * fy[r].  A thread computes
// its Gaussian's 16 column factors once and one row factor per row (32 MUFU.EX2 per tile instead
// of 256), and the sums factor too:
//   T(r,c) = gA(r,c).colour + gW(r,c) (+ gD z)
//   dColour = sum_r fy[r] * sum_c fx[c] gA(r,c)          S = sum_r fy[r] * sum_c fx[c] T(r,c)
//   column sums (for d/dpx, d/dsx): colS[c] = fx[c] * sum_r fy[r] T(r,c)
// => 8 FFMA per pixel-pair, issued as 4 packed f32x2 instructions + 1 LDS.128 per 4 pairs/channel.
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
            const float* __restrict__ g_alpha, const float* __restrict__ g_depth, float* __restrict__ gbuf) {
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
  // tile-major, channel planes: gbuf[tile][ch][256], ch = gA.r, gA.g, gA.b, gW, gD
  float* dst = gbuf + (size_t)tile * 5 * TILE_PIX + q;
  dst[0] = g.x;
  dst[TILE_PIX] = g.y;
  dst[2 * TILE_PIX] = g.z;
  dst[3 * TILE_PIX] = g.w;
  if (DEPTH) dst[4 * TILE_PIX] = gd;
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

__device__ __forceinline__ float2 bc2(float v) { return make_float2(v, v); }

template <bool DEPTH>
__global__ void __launch_bounds__(BB_THREADS)
blend_wsum_bwd_kernel(const ViewParams vp, const float4* __restrict__ rec, const int* __restrict__ vals,
                      const int2* __restrict__ ranges, const int* __restrict__ unit_start,
                      const int2* __restrict__ units, const float* __restrict__ gbuf, float* __restrict__ gacc) {
  __shared__ __align__(128) float sG[5][TILE_PIX];   // planes gA.r, gA.g, gA.b, gW, gD
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
    constexpr unsigned bytes = TILE_PIX * 4 * (DEPTH ? 5 : 4);
    mbar_expect_tx(&bar, bytes);
    bulk_g2s(&sG[0][0], gbuf + (size_t)tile * 5 * TILE_PIX, bytes, &bar);   // ONE TMA bulk copy per unit
  }
  const float x0 = tx * TILE + 0.5f, y0 = ty * TILE + 0.5f;
  bool waited = false;
  for (int base = 0; base < n; base += BB_THREADS) {
    const int i = base + threadIdx.x;
    const bool active = i < n;
    int id = 0;
    float4 a = make_float4(0.f, 0.f, -INFINITY, 0.f), b = make_float4(0.f, 0.f, 0.f, 0.f),
           col = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) {     // the record gather overlaps the bulk copy
      id = __ldg(vals + start + i);
      a = __ldg(rec + 3 * (size_t)id);
      b = __ldg(rec + 3 * (size_t)id + 1);
      col = __ldg(rec + 3 * (size_t)id + 2);
    }
    if (!waited) {
      mbar_wait(&bar, 0);
      waited = true;
    }
    // op == 0 (log2 op = -inf): the forward weight is 0 but torch's clamp_min(0) still passes
    // dL/dop = sum E*t at 0, so sweep with E instead of w and keep only S.
    const bool zero_op = (a.z == -INFINITY);
    const float lop = zero_op ? 0.0f : a.z;
    const float dx0 = x0 - a.x;
    float2 fx2[TILE / 2], Tc2[TILE / 2];
#pragma unroll
    for (int c2 = 0; c2 < TILE / 2; ++c2) {
      const float da = dx0 + (float)(2 * c2), db = da + 1.0f;
      fx2[c2] = make_float2(ex2_approx(fmaf(a.y * da, da, lop)), ex2_approx(fmaf(a.y * db, db, lop)));
      Tc2[c2] = make_float2(0.f, 0.f);
    }
    const float2 cr2 = bc2(col.x), cg2 = bc2(col.y), cb2 = bc2(col.z), z2 = bc2(col.w);
    float dR = 0.f, dG = 0.f, dB = 0.f, dZ = 0.f, S = 0.f, Sy = 0.f, Syy = 0.f;
#pragma unroll 1
    for (int r = 0; r < TILE; ++r) {
      const float dy = (y0 + (float)r) - b.x;
      const float fy = ex2_approx(b.y * dy * dy);
      const float2 fy2 = bc2(fy);
      float2 IR2 = make_float2(0.f, 0.f), IG2 = IR2, IB2 = IR2, ID2 = IR2, TR2 = IR2;
#pragma unroll
      for (int c4 = 0; c4 < TILE / 4; ++c4) {
        const float4 gr4 = *reinterpret_cast<const float4*>(&sG[0][r * TILE + 4 * c4]);
        const float4 gg4 = *reinterpret_cast<const float4*>(&sG[1][r * TILE + 4 * c4]);
        const float4 gb4 = *reinterpret_cast<const float4*>(&sG[2][r * TILE + 4 * c4]);
        const float4 gw4 = *reinterpret_cast<const float4*>(&sG[3][r * TILE + 4 * c4]);
        float4 gd4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (DEPTH) gd4 = *reinterpret_cast<const float4*>(&sG[4][r * TILE + 4 * c4]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c2 = 2 * c4 + h;
          const float2 gr2 = h ? make_float2(gr4.z, gr4.w) : make_float2(gr4.x, gr4.y);
          const float2 gg2 = h ? make_float2(gg4.z, gg4.w) : make_float2(gg4.x, gg4.y);
          const float2 gb2 = h ? make_float2(gb4.z, gb4.w) : make_float2(gb4.x, gb4.y);
          const float2 gw2 = h ? make_float2(gw4.z, gw4.w) : make_float2(gw4.x, gw4.y);
          float2 t2 = __ffma2_rn(gr2, cr2, __ffma2_rn(gg2, cg2, __ffma2_rn(gb2, cb2, gw2)));
          if (DEPTH) {
            const float2 gd2 = h ? make_float2(gd4.z, gd4.w) : make_float2(gd4.x, gd4.y);
            t2 = __ffma2_rn(gd2, z2, t2);
            ID2 = __ffma2_rn(fx2[c2], gd2, ID2);
          }
          Tc2[c2] = __ffma2_rn(fy2, t2, Tc2[c2]);
          TR2 = __ffma2_rn(fx2[c2], t2, TR2);
          IR2 = __ffma2_rn(fx2[c2], gr2, IR2);
          IG2 = __ffma2_rn(fx2[c2], gg2, IG2);
          IB2 = __ffma2_rn(fx2[c2], gb2, IB2);
        }
      }
      const float rowS = fy * (TR2.x + TR2.y);
      S += rowS;
      const float rd = rowS * dy;
      Sy += rd;
      Syy = fmaf(rd, dy, Syy);
      dR = fmaf(fy, IR2.x + IR2.y, dR);
      dG = fmaf(fy, IG2.x + IG2.y, dG);
      dB = fmaf(fy, IB2.x + IB2.y, dB);
      if (DEPTH) dZ = fmaf(fy, ID2.x + ID2.y, dZ);
    }
    float Sx = 0.f, Sxx = 0.f;
#pragma unroll
    for (int c2 = 0; c2 < TILE / 2; ++c2) {
      const float da = dx0 + (float)(2 * c2), db = da + 1.0f;
      const float ca = fx2[c2].x * Tc2[c2].x * da, cb = fx2[c2].y * Tc2[c2].y * db;
      Sx += ca + cb;
      Sxx = fmaf(ca, da, fmaf(cb, db, Sxx));
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
  if (g_depth != nullptr) {
    gbuf_kernel<true><<<vp.n_tiles, TILE_PIX, 0, st>>>(vp, acc, g_rgb, g_alpha, g_depth, gbuf);
    B2S_LAUNCH_CHECK();
    blend_wsum_bwd_kernel<true><<<(int)unit_cap, BB_THREADS, 0, st>>>(vp, rec, vals, ranges, unit_start, units, gbuf, gacc);
  } else {
    gbuf_kernel<false><<<vp.n_tiles, TILE_PIX, 0, st>>>(vp, acc, g_rgb, g_alpha, g_depth, gbuf);
    B2S_LAUNCH_CHECK();
    blend_wsum_bwd_kernel<false><<<(int)unit_cap, BB_THREADS, 0, st>>>(vp, rec, vals, ranges, unit_start, units, gbuf, gacc);
  }
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

}  // namespace b2s
