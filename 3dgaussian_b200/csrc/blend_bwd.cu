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
#include <cuda_fp16.h>
#include <limits.h>
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

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
    const float o0 = (view_bg(vp, 0) + A0) * inv, o1 = (view_bg(vp, 1) + A1) * inv, o2 = (view_bg(vp, 2) + A2) * inv;
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

// ---- per-pixel g-buffer, emitted directly as tensor-core B fragments (v5) -------------------------
// Same per-pixel numbers as gbuf_kernel, but stored for blend_wsum_bwd_mma_kernel: every plane value
// is scaled by the tile's power of two 2^sG (so that max|G| lands in [2^11, 2^12)), split into
// fp16 hi + fp16 lo (22 significant bits together) and written at the position the consumer's
// lane/register expects for the two operand orientations:
//   B1: N = row, K = column   (U = G . fx)        B2: N = column, K = row   (V = fy . G)
// m16n8k16 B fragment: lane (g = lane/4, t = lane%4) holds b0 = {K=2t, 2t+1}, b1 = {K=2t+8, 2t+9} at N = g.
// Register index = ((ch*2 + h)*2 + part)*2 + slot, h = N/8, part = hi|lo, slot = K/8; B2 follows B1.
// Layout in memory: [tile][register/4][lane] uint4  -> the consumer's loads are coalesced LDG.128.
// tile_scale[tile] = 2^-(sG + 16): undoes 2^sG and the 2^8 applied to each of fx, fy ... see the consumer.
// LOSS = true is the fit-loop fusion (python/fit_multiview_stub.py:292-297): the per-view loss
//   mean|rgb - tgt| + w_sil * mean|alpha - mask|
// is evaluated here straight from the accumulators (rgb and alpha are recomputed, not read back from images) and
// its image gradients feed the g-buffer in registers; the block's loss share goes to *loss_accum.
// The fragment image of the tile (NREG x 32 words) is assembled in shared memory and written with coalesced
// 16-byte stores.
// UMMA = true stores the same hi/lo planes as the four K-major shared-memory operand matrices of the tcgen05
// backward instead (see blend_wsum_bwd_umma_kernel): [U_hi | U_lo | V_hi | V_lo], each N = CH*16 rows x K = 16
// halves in the canonical no-swizzle layout (8-row x 16-byte core matrices),
//   U: row n = ch*16 + r, k = c        V: row n = ch*16 + c, k = r.
// With DEPTH the tcgen05 layout appends the gD plane as four small matrices (16 rows x 16 halves, 512 B each) behind
// the 8 KB of the four-plane matrices:  UD_hi | UD_lo | VD_hi | VD_lo,  UD: row n = r, k = c;  VD: row n = c, k = r.
// LOSS + DEPTH adds the fit script's depth term (python/fit_multiview_stub.py:298-303)
//   w_d * mean | depth / (max depth + 1e-6) - d_gt |
// whose gradient needs three per-view numbers computed beforehand by depth_stats_kernel (dstats = {max depth M,
// number of pixels attaining it, sum_p sign_p depth_p}): torch's max() backward hands the gradient of M to the arg-max
// pixels in equal shares.
template <bool DEPTH, bool LOSS, bool UMMA>
__global__ void __launch_bounds__(TILE_PIX)
gbuf_frag_kernel(const ViewParams vp, const float* __restrict__ acc, const float* __restrict__ g_rgb,
                 const float* __restrict__ g_alpha, const float* __restrict__ g_depth, const float* __restrict__ tgt,
                 const float* __restrict__ mask, const float* __restrict__ depth_gt, const float* __restrict__ dstats,
                 float w_sil, float w_depth, float scale, float* __restrict__ loss_accum,
                 uint32_t* __restrict__ frag, float* __restrict__ tile_scale, const uint8_t* __restrict__ tgt8,
                 const uint8_t* __restrict__ mask8) {
  constexpr int CH = DEPTH ? 5 : 4;
  constexpr int NREG = CH * 16;
  __shared__ float wmax[TILE_PIX / 32];
  __shared__ float wloss[TILE_PIX / 32];
  __shared__ __align__(16) __half sfrag[NREG * 32 * 2];
  const int tile = blockIdx.x, q = threadIdx.x;
  const int r = q >> 4, c = q & 15;
  const int xi = (tile % vp.tiles_x) * TILE + c, yi = (tile / vp.tiles_x) * TILE + r;
  const size_t hw = (size_t)vp.width * vp.height;
  float v[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  float loss = 0.f;
  if (xi < vp.width && yi < vp.height) {
    const size_t p = (size_t)yi * vp.width + xi;
    const float A0 = acc[p], A1 = acc[hw + p], A2 = acc[2 * hw + p], W = acc[3 * hw + p];
    const float inv = 1.0f / (1.0f + W);
    const float o0 = (view_bg(vp, 0) + A0) * inv, o1 = (view_bg(vp, 1) + A1) * inv, o2 = (view_bg(vp, 2) + A2) * inv;
    float gr0, gr1, gr2, ga = 0.f;
    if constexpr (LOSS) {
      const float inv3 = 1.0f / (3.0f * (float)hw), inv1 = 1.0f / (float)hw;
      float t0, t1, t2;
      if (tgt8 != nullptr) {          // image bytes: the same value / 255 a host-side conversion would have produced
        t0 = (float)tgt8[3 * p] / 255.0f; t1 = (float)tgt8[3 * p + 1] / 255.0f; t2 = (float)tgt8[3 * p + 2] / 255.0f;
      } else {
        t0 = tgt[3 * p]; t1 = tgt[3 * p + 1]; t2 = tgt[3 * p + 2];
      }
      const float d0 = fminf(fmaxf(o0, 0.f), 1.f) - t0, d1 = fminf(fmaxf(o1, 0.f), 1.f) - t1,
                  d2 = fminf(fmaxf(o2, 0.f), 1.f) - t2;
      loss = (fabsf(d0) + fabsf(d1) + fabsf(d2)) * inv3;
      const float k = scale * inv3;
      gr0 = k * (float)((d0 > 0.f) - (d0 < 0.f));
      gr1 = k * (float)((d1 > 0.f) - (d1 < 0.f));
      gr2 = k * (float)((d2 > 0.f) - (d2 < 0.f));
      if (mask != nullptr || mask8 != nullptr) {
        const float mk = mask8 != nullptr ? (float)mask8[p] / 255.0f : mask[p];
        const float da = fminf(fmaxf(W * inv, 0.f), 1.f) - mk;
        loss = fmaf(w_sil * inv1, fabsf(da), loss);
        ga = scale * w_sil * inv1 * (float)((da > 0.f) - (da < 0.f));
      }
    } else {
      gr0 = g_rgb[3 * p]; gr1 = g_rgb[3 * p + 1]; gr2 = g_rgb[3 * p + 2];
      if (g_alpha != nullptr) ga = g_alpha[p];
    }
    v[0] = (o0 >= 0.f && o0 <= 1.f) ? gr0 * inv : 0.f;
    v[1] = (o1 >= 0.f && o1 <= 1.f) ? gr1 * inv : 0.f;
    v[2] = (o2 >= 0.f && o2 <= 1.f) ? gr2 * inv : 0.f;
    v[3] = -(v[0] * o0 + v[1] * o1 + v[2] * o2);
    v[3] = fmaf(ga, inv * inv, v[3]);
    if (DEPTH) {
      const float D = acc[4 * hw + p];
      const float iw = 1.0f / (W + 1e-6f);
      float gdep;
      if constexpr (LOSS) {
        // d_pred = depth / (M + 1e-6);  dL/ddepth_q = k (s_q / (M+eps) - [q in argmax] / cnt * sum_p s_p depth_p / (M+eps)^2)
        const float inv1 = 1.0f / (float)hw;
        const float dep = fmaxf(D / (W + 1e-6f), 0.0f);     // the very expression of depth_max_kernel: dep == M must be exact
        const float M = dstats[0], cnt = dstats[1], ssum = dstats[2];
        const float im = 1.0f / (M + 1e-6f);
        const float dd = dep * im - depth_gt[p];
        const float sg = (float)((dd > 0.f) - (dd < 0.f));
        loss = fmaf(w_depth * inv1, fabsf(dd), loss);
        const float k = scale * w_depth * inv1;
        gdep = k * sg * im;
        if (dep == M && cnt > 0.0f) gdep -= k * ssum * im * im / cnt;
        gdep = (D * iw >= 0.f) ? gdep : 0.f;
      } else {
        gdep = (D * iw >= 0.f) ? g_depth[p] : 0.f;
      }
      v[4] = gdep * iw;
      v[3] = fmaf(-gdep * D, iw * iw, v[3]);
    }
  }
  float m = 0.f;
#pragma unroll
  for (int ch = 0; ch < CH; ++ch) m = fmaxf(m, fabsf(v[ch]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (LOSS) loss += __shfl_xor_sync(0xffffffffu, loss, o);
  }
  if ((q & 31) == 0) { wmax[q >> 5] = m; wloss[q >> 5] = loss; }
  __syncthreads();
  m = 0.f;
#pragma unroll
  for (int w = 0; w < TILE_PIX / 32; ++w) m = fmaxf(m, wmax[w]);
  int sG = 0;
  if (m > 0.f && m < INFINITY) {
    int e;
    frexpf(m, &e);                 // 2^(e-1) <= m < 2^e
    sG = min(max(12 - e, -60), 60);
  }
  const float sc = ldexpf(1.0f, sG);
  if (q == 0) {
    tile_scale[tile] = ldexpf(1.0f, -(sG + 16));
    if (LOSS) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < TILE_PIX / 32; ++w) t += wloss[w];
      atomicAdd(loss_accum, scale * t);
    }
  }
#pragma unroll
  for (int ch = 0; ch < CH; ++ch) {
    const float x = v[ch] * sc;
    const __half hi = __float2half_rn(x);
    const __half lo = __float2half_rn(x - __half2float(hi));
#pragma unroll
    for (int part = 0; part < 2; ++part) {
      const __half val = part ? lo : hi;
      if constexpr (UMMA) {
        if (ch < 4) {
          constexpr int NR = 64;                    // operand rows of the four-plane matrices; one matrix = NR x 16 halves
          // operand row (= TMEM column of the accumulator) = quad * 16 + plane * 4 + index in quad: the epilogue's
          // unit of work -- 4 rows / columns x 4 planes -- is 16 consecutive columns, ONE tcgen05.ld
          const int nu = (r >> 2) * 16 + ch * 4 + (r & 3), nv = (c >> 2) * 16 + ch * 4 + (c & 3);
          sfrag[(0 + part) * NR * 16 + (c >> 3) * NR * 8 + (nu >> 3) * 64 + (nu & 7) * 8 + (c & 7)] = val;   // U: k = c
          sfrag[(2 + part) * NR * 16 + (r >> 3) * NR * 8 + (nv >> 3) * 64 + (nv & 7) * 8 + (r & 7)] = val;   // V: k = r
        } else {                                    // gD: 16-row matrices behind the 4 x 2 KB block
          sfrag[4096 + (0 + part) * 256 + (c >> 3) * 128 + (r >> 3) * 64 + (r & 7) * 8 + (c & 7)] = val;     // UD: n = r, k = c
          sfrag[4096 + (2 + part) * 256 + (r >> 3) * 128 + (c >> 3) * 64 + (c & 7) * 8 + (r & 7)] = val;     // VD: n = c, k = r
        }
      } else {
        {   // B1: N = row r, K = column c
          const int h = r >> 3, g = r & 7, slot = c >> 3, t = (c & 7) >> 1, half = c & 1;
          const int reg = ((ch * 2 + h) * 2 + part) * 2 + slot, lane = g * 4 + t;
          sfrag[((size_t)((reg >> 2) * 32 + lane) * 4 + (reg & 3)) * 2 + half] = val;
        }
        {   // B2: N = column c, K = row r
          const int h = c >> 3, g = c & 7, slot = r >> 3, t = (r & 7) >> 1, half = r & 1;
          const int reg = CH * 8 + ((ch * 2 + h) * 2 + part) * 2 + slot, lane = g * 4 + t;
          sfrag[((size_t)((reg >> 2) * 32 + lane) * 4 + (reg & 3)) * 2 + half] = val;
        }
      }
    }
  }
  __syncthreads();
  // tcgen05 layout: tiles are GBUF_FRAG_WORDS apart (10 KB, with or without the gD matrices); mma.sync fragments: packed
  uint4* out = reinterpret_cast<uint4*>(frag + (size_t)tile * (UMMA ? GBUF_FRAG_WORDS : NREG * 32));
  const uint4* src = reinterpret_cast<const uint4*>(sfrag);
#pragma unroll
  for (int k = q; k < NREG * 32 / 4; k += TILE_PIX) out[k] = src[k];
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
  const int start = rg.x + ud.y * vp.seg;
  const int n = min(vp.seg, rg.y - start);
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
      red_add_v4(dst, dR, dG, dB, Syy);
      red_add_v4(dst + 4, S, Sx, Sxx, Sy);
      if (DEPTH) atomicAdd(dst + 8, dZ);
    }
  }
}

// ---- v5: the separable sums as small GEMMs on the tensor cores -----------------------------------
// With w(r,c) = op fy[r] fx[c] the per-Gaussian sums of one tile are two matrix products over the
// 16x16 g-buffer planes G_ch (ch = gA.r, gA.g, gA.b, gW [, gD]):
//   U_ch[i][r] = sum_c fx_i[c] G_ch[r][c]      (contract the columns)
//   V_ch[i][c] = sum_r fy_i[r] G_ch[r][c]      (contract the rows)
// followed by O(16) FP32 work per Gaussian:
//   dColour_ch = sum_r fy[r] U_ch[r];  rowT[r] = colour . U[r] + U_W[r];  S,Sy,Syy = sum_r fy[r] dy^k rowT[r]
//   colT[c] = colour . V[c] + V_W[c];  Sx,Sxx = sum_c fx[c] dx^k colT[c]
// i.e. 2 x 16 x 16 x CH MACs per (Gaussian,tile) on the tensor pipe (mma.sync m16n8k16, FP32
// accumulate: K = 16 is exactly the tile edge) + ~70 FP32 ops, instead of 8 FP32 FMAs for each of
// the 256 pixels.
// Precision (gradients are held to relative L2 <= 1e-3 by north_star):
//   * the factors fx, fy enter as single fp16 (2^-11): their rounding is common to all planes, so it
//     commutes with the per-Gaussian plane combination colour.G + G_W and stays a 2^-11-relative
//     perturbation of each term;
//   * the planes enter as fp16 hi + fp16 lo (2 MMAs): rounding them independently (as the first
//     TF32 version did) is amplified by the cancellation inside colour.G + G_W = gA.(colour - out) and
//     measured 1.2e-3..2.4e-3 on the scale gradients; with hi + lo the planes carry 22 bits.
//   * range: planes are pre-scaled per tile by a power of two (gbuf_frag_kernel), factors by 2^8
//     (added to the exponent of ex2, free); opacity is applied in FP32 afterwards.
//
// One warp owns a work unit (tile, <= SEG Gaussians) and processes 16 Gaussians per step with
// M = Gaussians.  Lane (g = lane/4, t = lane%4) owns Gaussians g and g+8 of the step and the pixel
// coordinates idx(q) = 8*(q/2) + 2t + (q%2), q = 0..3, on both axes: the K slots of the A fragment a
// lane builds (its 4 fx / 4 fy values per Gaussian) are exactly the N slots (rows / columns) it
// receives back in the accumulator fragment.  The tile's planes live in registers as B fragments
// for the whole unit.  Records are staged per warp with cp.async, ids one stage ahead.
constexpr int BM_WARPS = 4;
constexpr int BM_STAGE = 32;      // Gaussians per cp.async stage = 2 MMA steps
constexpr int BM_STAGES = 3;

struct BmStage {
  float4 a[BM_STAGE];   // {px, qx, log2 op, bbox x}
  float4 b[BM_STAGE];   // {py, qy, 0, bbox y}
  float4 c[BM_STAGE];   // {r, g, b, zabs}
  int id[BM_STAGE];
};

__device__ __forceinline__ uint32_t pack_h2(float lo_k, float hi_k) {   // {K even, K odd} -> f16x2
  const __half2 h = __floats2half2_rn(lo_k, hi_k);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// D (16x8, fp32) = A (16x16 f16, row) * B (16x8 f16, col) + C
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, const float (&c)[4]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};\n"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
}

// One MMA step of the backward: 16 Gaussians (slots bt*16 .. bt*16+15 of the staged chunk) against the tile's
// planes.  getB(r4) returns fragment registers 4*r4 .. 4*r4+3 of this lane (see gbuf_frag_kernel).
template <bool DEPTH, class GetB>
__device__ __forceinline__ void bwd_mma_step(const BmStage& st, int bt, int g, int t, const float (&cx)[4],
                                             const float (&cy)[4], float k_us, GetB getB, float* __restrict__ gacc) {
  constexpr int CH = DEPTH ? 5 : 4;
  const float zero4[4] = {0.f, 0.f, 0.f, 0.f};
  const int j[2] = {bt * 16 + g, bt * 16 + g + 8};
  float4 ra[2], rb[2], rc[2];
  // factors scaled by 2^8 (fp16 range), WITHOUT opacity, and the pixel offsets, as (q = 2h, 2h+1) pairs:
  // the FP32 epilogue below runs on packed f32x2 instructions (half the issue slots)
  float2 fx2[2][2], fy2[2][2], dx2[2][2], dy2[2][2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    ra[e] = st.a[j[e]];
    rb[e] = st.b[j[e]];
    rc[e] = st.c[j[e]];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float dxa = cx[2 * h] - ra[e].x, dxb = cx[2 * h + 1] - ra[e].x;
      const float dya = cy[2 * h] - rb[e].x, dyb = cy[2 * h + 1] - rb[e].x;
      dx2[e][h] = make_float2(dxa, dxb);
      dy2[e][h] = make_float2(dya, dyb);
      fx2[e][h] = make_float2(ex2_approx(fmaf(ra[e].y * dxa, dxa, 8.0f)), ex2_approx(fmaf(ra[e].y * dxb, dxb, 8.0f)));
      fy2[e][h] = make_float2(ex2_approx(fmaf(rb[e].y * dya, dya, 8.0f)), ex2_approx(fmaf(rb[e].y * dyb, dyb, 8.0f)));
    }
  }
  // A fragments (m16n8k16): a0 = (Gaussian g; K = 2t, 2t+1), a1 = (g+8; same K), a2 = (g; K = 2t+8, 2t+9),
  // a3 = (g+8; same) -- K slots 2t, 2t+1, 2t+8, 2t+9 are this lane's coordinates q = 0..3.
  uint32_t Ax[4], Ay[4];
  Ax[0] = pack_h2(fx2[0][0].x, fx2[0][0].y); Ax[1] = pack_h2(fx2[1][0].x, fx2[1][0].y);
  Ax[2] = pack_h2(fx2[0][1].x, fx2[0][1].y); Ax[3] = pack_h2(fx2[1][1].x, fx2[1][1].y);
  Ay[0] = pack_h2(fy2[0][0].x, fy2[0][0].y); Ay[1] = pack_h2(fy2[1][0].x, fy2[1][0].y);
  Ay[2] = pack_h2(fy2[0][1].x, fy2[0][1].y); Ay[3] = pack_h2(fy2[1][1].x, fy2[1][1].y);
  // per-Gaussian partial sums of this lane as (q even, q odd) pairs, added horizontally after the sweeps:
  //   Q[e][0..3] = dR dG dB dZ     Q[e][4..7] = S Sx Sxx Sy     Q[e][8] = Syy        (e = 0 -> Gaussian g, 1 -> g+8)
  float2 Q[2][9];
#pragma unroll
  for (int e = 0; e < 2; ++e)
#pragma unroll
    for (int k = 0; k < 9; ++k) Q[e][k] = make_float2(0.f, 0.f);
  float2 col2[2][4];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    col2[e][0] = bc2(rc[e].x); col2[e][1] = bc2(rc[e].y); col2[e][2] = bc2(rc[e].z); col2[e][3] = bc2(rc[e].w);
  }

#pragma unroll
  for (int h = 0; h < 2; ++h) {
    // ---- U = fx . G for rows 8h..8h+7: accumulator pair (d[2e], d[2e+1]) = Gaussian e at rows q = 2h, 2h+1
    float D[CH][4];
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
      const uint4 b = getB(ch * 2 + h);              // {hi k0-7, hi k8-15, lo k0-7, lo k8-15}
      mma_f16(D[ch], Ax, b.z, b.w, zero4);
      mma_f16(D[ch], Ax, b.x, b.y, D[ch]);
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float2 uR = make_float2(D[0][2 * e], D[0][2 * e + 1]), uG = make_float2(D[1][2 * e], D[1][2 * e + 1]),
                   uB = make_float2(D[2][2 * e], D[2][2 * e + 1]), uW = make_float2(D[3][2 * e], D[3][2 * e + 1]);
      float2 T = __ffma2_rn(col2[e][0], uR, __ffma2_rn(col2[e][1], uG, __ffma2_rn(col2[e][2], uB, uW)));
      if (DEPTH) {
        const float2 uD = make_float2(D[CH - 1][2 * e], D[CH - 1][2 * e + 1]);
        T = __ffma2_rn(col2[e][3], uD, T);
        Q[e][3] = __ffma2_rn(fy2[e][h], uD, Q[e][3]);
      }
      const float2 a = __fmul2_rn(fy2[e][h], T);
      Q[e][4] = __fadd2_rn(Q[e][4], a);
      Q[e][7] = __ffma2_rn(a, dy2[e][h], Q[e][7]);
      Q[e][8] = __ffma2_rn(__fmul2_rn(a, dy2[e][h]), dy2[e][h], Q[e][8]);
      Q[e][0] = __ffma2_rn(fy2[e][h], uR, Q[e][0]);
      Q[e][1] = __ffma2_rn(fy2[e][h], uG, Q[e][1]);
      Q[e][2] = __ffma2_rn(fy2[e][h], uB, Q[e][2]);
    }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    // ---- V = fy . G for columns 8h..8h+7
    float D[CH][4];
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
      const uint4 b = getB(CH * 2 + ch * 2 + h);
      mma_f16(D[ch], Ay, b.z, b.w, zero4);
      mma_f16(D[ch], Ay, b.x, b.y, D[ch]);
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float2 vR = make_float2(D[0][2 * e], D[0][2 * e + 1]), vG = make_float2(D[1][2 * e], D[1][2 * e + 1]),
                   vB = make_float2(D[2][2 * e], D[2][2 * e + 1]), vW = make_float2(D[3][2 * e], D[3][2 * e + 1]);
      float2 T = __ffma2_rn(col2[e][0], vR, __ffma2_rn(col2[e][1], vG, __ffma2_rn(col2[e][2], vB, vW)));
      if (DEPTH) T = __ffma2_rn(col2[e][3], make_float2(D[CH - 1][2 * e], D[CH - 1][2 * e + 1]), T);
      const float2 b = __fmul2_rn(__fmul2_rn(fx2[e][h], T), dx2[e][h]);
      Q[e][5] = __fadd2_rn(Q[e][5], b);
      Q[e][6] = __ffma2_rn(b, dx2[e][h], Q[e][6]);
    }
  }
  float P[2][8], Syy[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
#pragma unroll
    for (int k = 0; k < 8; ++k) P[e][k] = Q[e][k].x + Q[e][k].y;
    Syy[e] = Q[e][8].x + Q[e][8].y;
  }
  // ---- sum over the quad (the 4 lanes t = 0..3 share Gaussians g, g+8) by transposition: lane t ends
  // with group t of {e0: dR dG dB dZ | e0: S Sx Sxx Sy | e1: dR.. | e1: S..} summed over the quad.
  const bool up = (t & 2) != 0;                // lanes 2,3 keep Gaussian g+8, lanes 0,1 keep Gaussian g
  const bool odd = (t & 1) != 0;               // odd lanes keep {S Sx Sxx Sy}, even lanes {dR dG dB dZ}
  float v4[4];
  {
    float keep[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float send = up ? P[0][k] : P[1][k];
      keep[k] = (up ? P[1][k] : P[0][k]) + __shfl_xor_sync(0xffffffffu, send, 2);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float send = odd ? keep[k] : keep[4 + k];
      v4[k] = (odd ? keep[4 + k] : keep[k]) + __shfl_xor_sync(0xffffffffu, send, 1);
    }
  }
  float syy = (up ? Syy[1] : Syy[0]) + __shfl_xor_sync(0xffffffffu, up ? Syy[0] : Syy[1], 2);
  syy += __shfl_xor_sync(0xffffffffu, syy, 1);
  {
    const int id = st.id[up ? j[1] : j[0]];
    const float lop = up ? ra[1].z : ra[0].z;
    // op == 0 (log2 op = -inf): forward weight 0, but clamp_min(0) passes dL/dop = sum E*t: keep S with op = 1
    const bool zop = (lop == -INFINITY);
    const float opk = zop ? k_us : ex2_approx(lop) * k_us;   // opacity and the 2^-(sG+16) un-scaling, once
    if (id >= 0) {
      float* dst = gacc + (size_t)id * GACC_F;
      if (odd) {
        red_add_v4(dst + 4, v4[0] * opk, zop ? 0.f : v4[1] * opk, zop ? 0.f : v4[2] * opk, zop ? 0.f : v4[3] * opk);
      } else if (!zop) {
        red_add_v4(dst, v4[0] * opk, v4[1] * opk, v4[2] * opk, syy * opk);
        if (DEPTH) atomicAdd(dst + 8, v4[3] * opk);
      }
    }
  }
}

template <bool DEPTH, int MINB>
__global__ void __launch_bounds__(BM_WARPS * 32, MINB)
blend_wsum_bwd_mma_kernel(const ViewParams vp, const float4* __restrict__ rec, const int* __restrict__ vals,
                          const int2* __restrict__ ranges, const int* __restrict__ unit_start,
                          const int2* __restrict__ units, const uint4* __restrict__ frag,
                          const float* __restrict__ tile_scale, float* __restrict__ gacc) {
  constexpr int CH = DEPTH ? 5 : 4;
  constexpr int NREG = CH * 16;
  __shared__ __align__(16) BmStage ring[BM_WARPS][BM_STAGES];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u = blockIdx.x * BM_WARPS + warp;
  if (u >= unit_start[vp.n_tiles]) return;             // warps are independent: no block barrier below
  const int2 ud = units[u];
  const int tile = ud.x;
  const int2 rg = ranges[tile];
  const int start = rg.x + ud.y * vp.seg;
  const int n = min(vp.seg, rg.y - start);
  if (n <= 0) return;
  const int tx = tile % vp.tiles_x, ty = tile / vp.tiles_x;
  const int g = lane >> 2, t = lane & 3;
  BmStage* my = ring[warp];

  // ---- stage the first records while the plane fragments are fetched
  const int nchunks = (n + BM_STAGE - 1) / BM_STAGE;
  auto issue = [&](int c, int id) {      // chunk c, this lane's Gaussian id (already loaded)
    if (c < nchunks) {
      BmStage& st = my[c % BM_STAGES];
      if (c * BM_STAGE + lane < n) {
        const float4* src = rec + 3 * (size_t)id;
        cp_async16_b(&st.a[lane], src);
        cp_async16_b(&st.b[lane], src + 1);
        cp_async16_b(&st.c[lane], src + 2);
        st.id[lane] = id;
      } else {   // padding of the last step: a record whose factors underflow to exactly 0 (no selects below)
        st.a[lane] = make_float4(1e18f, -1.0f, 0.0f, 0.0f);
        st.b[lane] = make_float4(1e18f, -1.0f, 0.0f, 0.0f);
        st.c[lane] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        st.id[lane] = -1;
      }
    }
    cp_async_commit_b();
  };
  auto load_id = [&](int c) -> int {
    const int i = c * BM_STAGE + lane;
    return (c < nchunks && i < n) ? __ldg(vals + start + i) : 0;
  };
  {
    int ids[BM_STAGES - 1];
#pragma unroll
    for (int c = 0; c < BM_STAGES - 1; ++c) ids[c] = load_id(c);
#pragma unroll
    for (int c = 0; c < BM_STAGES - 1; ++c) issue(c, ids[c]);
  }
  int id_pf = load_id(BM_STAGES - 1);

  // ---- the tile's planes as B fragments (fp16 hi/lo), kept for the whole unit; see gbuf_frag_kernel
  uint32_t B[NREG];
  {
    const uint4* ft = frag + (size_t)tile * (NREG / 4) * 32 + lane;
#pragma unroll
    for (int r4 = 0; r4 < NREG / 4; ++r4) {
      const uint4 v = __ldg(ft + r4 * 32);
      B[4 * r4] = v.x; B[4 * r4 + 1] = v.y; B[4 * r4 + 2] = v.z; B[4 * r4 + 3] = v.w;
    }
  }
  const float k_us = __ldg(tile_scale + tile);       // 2^-(sG+16): planes' 2^sG and the 2^8 of each factor
  // this lane's four pixel-centre coordinates on each axis
  float cx[4], cy[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int idx = 8 * (q >> 1) + 2 * t + (q & 1);
    cx[q] = (float)(tx * TILE + idx) + 0.5f;
    cy[q] = (float)(ty * TILE + idx) + 0.5f;
  }

  for (int c = 0; c < nchunks; ++c) {
    issue(c + BM_STAGES - 1, id_pf);
    id_pf = load_id(c + BM_STAGES);
    cp_async_wait_b<BM_STAGES - 1>();
    __syncwarp();
    const BmStage& st = my[c % BM_STAGES];
#pragma unroll 1
    for (int bt = 0; bt < 2; ++bt) {
      const int base = c * BM_STAGE + bt * 16;
      if (base >= n) break;                        // warp-uniform
      bwd_mma_step<DEPTH>(st, bt, g, t, cx, cy, k_us,
                          [&](int r4) { return make_uint4(B[4 * r4], B[4 * r4 + 1], B[4 * r4 + 2], B[4 * r4 + 3]); }, gacc);
    }
    __syncwarp();
  }
  cp_async_wait_b<0>();
}

// ---- v7: the separable sums on the 5th-generation tensor cores (tcgen05 / TMEM) ---------------------------
// The backward's natural M dimension is the GAUSSIAN: for the 128 Gaussians of a batch
//   U[i][(ch,r)] = sum_c fx_i[c] G_ch[r][c]      V[i][(ch,c)] = sum_r fy_i[r] G_ch[r][c]
// are two 128 x 64 x 16 products -- one tcgen05.mma each per plane half (hi, lo accumulate in TMEM) -- against
// the tile's planes, which sit in shared memory as K-major operand matrices for the whole unit.  One thread owns
// one Gaussian: it evaluates its 16 + 16 factors (by recurrence, 14 MUFU.EX2), writes them as one fp16 row of the two A operands,
// and after the MMA reads ITS accumulator row (TMEM lane = thread) with tcgen05.ld: all 16 rows and 16 columns
// of its Gaussian arrive in its own registers, so the FP32 epilogue is thread local -- no quad shuffles, no
// fragment bookkeeping, and no HMMA issue slots (mma.sync tops out at a quarter of the tcgen05 rate).
//   CTA = 128 threads = 4 warps (warp w reads TMEM lanes 32w..32w+31), 128 TMEM columns (U: 0..63, V: 64..127),
//   28 KB of shared memory; 4 CTAs per SM cover each other's MMA round trips.
// Operand layout: K-major, no swizzle (8-row x 16-byte core matrices; row groups 128 B apart, the two K chunks
// `rows*16` B apart), validated by profiles/microbench/umma_probe.cu.
constexpr int BT_THREADS = 128;

// one quad: 16 consecutive accumulator columns = [plane][index in quad]
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[4][4]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=f"(v[0][0]), "=f"(v[0][1]), "=f"(v[0][2]), "=f"(v[0][3]), "=f"(v[1][0]), "=f"(v[1][1]), "=f"(v[1][2]), "=f"(v[1][3]),
        "=f"(v[2][0]), "=f"(v[2][1]), "=f"(v[2][2]), "=f"(v[2][3]), "=f"(v[3][0]), "=f"(v[3][1]), "=f"(v[3][2]), "=f"(v[3][3])
      : "r"(taddr));
}
// tcgen05.wait::ld for the 16 registers of one quad; they are listed as in/out operands so that no use of them can
// be scheduled above the wait
__device__ __forceinline__ void tmem_wait_quad(float (&q)[4][4]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;\n"
               : "+f"(q[0][0]), "+f"(q[0][1]), "+f"(q[0][2]), "+f"(q[0][3]), "+f"(q[1][0]), "+f"(q[1][1]), "+f"(q[1][2]),
                 "+f"(q[1][3]), "+f"(q[2][0]), "+f"(q[2][1]), "+f"(q[2][2]), "+f"(q[2][3]), "+f"(q[3][0]), "+f"(q[3][1]),
                 "+f"(q[3][2]), "+f"(q[3][3])
               :
               : "memory");
}

// Thread-local epilogue of the tcgen05 backward: this thread's accumulator row -- U at columns [0,64) (quad*16 +
// plane*4 + row%4), V at [64,128) (same with columns) of `taddr` -- is streamed through registers in 8 quads of 4 rows /
// columns x 4 planes.  The loads are software pipelined: quad q+1 is requested right after quad q has landed, so its
// TMEM round trip hides behind the FP32 work of quad q (tcgen05.wait::ld waits for ALL outstanding loads, so at most
// one quad may be in flight when it is issued).  fy_pair(p) / fx_pair(p): factors of rows / columns 2p, 2p+1.
struct BwdSums {
  float2 aR, aG, aB, aS, aSy, aSyy, aSx, aSxx, aZ;
};
// four consecutive accumulator columns (the gD plane's share of a quad)
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];\n"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_quad5(float (&q)[4][4], float (&d)[4]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;\n"
               : "+f"(q[0][0]), "+f"(q[0][1]), "+f"(q[0][2]), "+f"(q[0][3]), "+f"(q[1][0]), "+f"(q[1][1]), "+f"(q[1][2]),
                 "+f"(q[1][3]), "+f"(q[2][0]), "+f"(q[2][1]), "+f"(q[2][2]), "+f"(q[2][3]), "+f"(q[3][0]), "+f"(q[3][1]),
                 "+f"(q[3][2]), "+f"(q[3][3]), "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               :
               : "memory");
}
// DEPTH: the gD plane's sums U_D (rows) / V_D (columns) sit in a second TMEM region `taddr_d` (16 + 16 columns); they
// enter T = colour . v + v_W + z v_D and, along the rows, the depth gradient dZ = sum_r fy[r] U_D[r].
template <bool DEPTH, class FyPair, class FxPair>
__device__ __forceinline__ void umma_epilogue(uint32_t taddr, uint32_t taddr_d, float dx0, float dy0, float2 cR, float2 cG,
                                              float2 cB, float2 cZ, FyPair fy_pair, FxPair fx_pair, BwdSums& A) {
  float buf[2][4][4];
  float bd[2][4];
  auto request = [&](int q, float (&dst)[4][4], float (&dd)[4]) {
    tmem_ld16(taddr + (uint32_t)(q * 16), dst);
    if constexpr (DEPTH) tmem_ld4(taddr_d + (uint32_t)(q * 4), dd);
  };
  A.aR = A.aG = A.aB = A.aS = A.aSy = A.aSyy = A.aSx = A.aSxx = A.aZ = make_float2(0.f, 0.f);
  request(0, buf[0], bd[0]);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    if constexpr (DEPTH) tmem_wait_quad5(buf[q & 1], bd[q & 1]); else tmem_wait_quad(buf[q & 1]);
    if (q < 7) request(q + 1, buf[(q + 1) & 1], bd[(q + 1) & 1]);
    float (&b)[4][4] = buf[q & 1];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int i0 = 4 * (q & 3) + 2 * j;                  // row (U) or column (V) of the pair
      const float2 vR = make_float2(b[0][2 * j], b[0][2 * j + 1]), vG = make_float2(b[1][2 * j], b[1][2 * j + 1]),
                   vB = make_float2(b[2][2 * j], b[2][2 * j + 1]), vW = make_float2(b[3][2 * j], b[3][2 * j + 1]);
      float2 T = __ffma2_rn(cR, vR, __ffma2_rn(cG, vG, __ffma2_rn(cB, vB, vW)));
      if constexpr (DEPTH) {
        const float2 vD = make_float2(bd[q & 1][2 * j], bd[q & 1][2 * j + 1]);
        T = __ffma2_rn(cZ, vD, T);
        if (q < 4) A.aZ = __ffma2_rn(fy_pair(i0 >> 1), vD, A.aZ);
      }
      if (q < 4) {
        const float2 f = fy_pair(i0 >> 1);
        const float2 dy = make_float2(dy0 + (float)i0, dy0 + (float)(i0 + 1));
        const float2 a = __fmul2_rn(f, T);
        A.aS = __fadd2_rn(A.aS, a);
        A.aSy = __ffma2_rn(a, dy, A.aSy);
        A.aSyy = __ffma2_rn(__fmul2_rn(a, dy), dy, A.aSyy);
        A.aR = __ffma2_rn(f, vR, A.aR);
        A.aG = __ffma2_rn(f, vG, A.aG);
        A.aB = __ffma2_rn(f, vB, A.aB);
      } else {
        const float2 dx = make_float2(dx0 + (float)i0, dx0 + (float)(i0 + 1));
        const float2 bb = __fmul2_rn(__fmul2_rn(fx_pair(i0 >> 1), T), dx);
        A.aSx = __fadd2_rn(A.aSx, bb);
        A.aSxx = __ffma2_rn(bb, dx, A.aSxx);
      }
    }
  }
}

// DEPTH (a depth gradient is present): the gD plane adds four N = 16 products per step into a SECOND TMEM allocation of
// 32 columns (U_D | V_D) -- 128 + 32 columns per CTA, three CTAs per SM -- read in the epilogue next to each quad; the
// thread's sums gain dZ, which leaves as a third, scalar RED.
template <bool RECUR, bool DEPTH>
__global__ void __launch_bounds__(BT_THREADS, DEPTH ? 3 : 4)
blend_wsum_bwd_umma_kernel(const ViewParams vp, const float4* __restrict__ rec, const int* __restrict__ vals,
                           const int4* __restrict__ udesc, const Counters* __restrict__ counters,
                           const uint4* __restrict__ planes,
                           const float* __restrict__ tile_scale, float* __restrict__ gacc) {
  constexpr int NR = 64;                                   // operand rows of a plane matrix: 4 planes x 16
  constexpr uint32_t PLANE4_BYTES = 4 * NR * 16 * 2;       // U_hi | U_lo | V_hi | V_lo, 2 KB each
  constexpr uint32_t PLANE_BYTES = PLANE4_BYTES + (DEPTH ? 4 * 16 * 16 * 2 : 0);   // + UD_hi | UD_lo | VD_hi | VD_lo, 512 B each
  constexpr uint32_t PLANE_STRIDE = GBUF_FRAG_WORDS * 4;   // bytes between tiles in global memory (10 KB: room for both layouts)
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(NR >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // f32 += f16 x f16, K-major
  constexpr uint32_t IDESC_D = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // N = 16: the gD plane
  __shared__ __align__(128) uint4 sP[2][PLANE_BYTES / 16];       // the planes of the current and of the next unit's tile
  __shared__ __align__(128) uint4 sA[2][128 * 16 * 2 / 16];      // A operands: fx rows, fy rows (4 KB each)
  __shared__ __align__(16) float4 sRec[DEPTH ? 3 : 2][BT_THREADS];   // x / y [/ colour + zabs] record of the next step, one private slot per thread
  __shared__ __align__(8) unsigned long long bar_load[2], bar_mma;
  __shared__ uint32_t tmem_base_s, tmem_base_d;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int nunits = counters->n_ne;                       // entries of the unit descriptor table (non-empty units, largest first)
  if ((int)blockIdx.x >= nunits) return;                   // block-uniform

  if (tid == 0) {
    mbar_init(&bar_load[0], 1);
    mbar_init(&bar_load[1], 1);
    mbar_init(&bar_mma, 1);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(128) : "memory");
    if constexpr (DEPTH)
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_d)), "r"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t tmem_d = DEPTH ? tmem_base_d : 0u;
  const uint32_t taddr_d = tmem_d + ((uint32_t)(warp * 32) << 16);
  // this thread's A rows (row tid): K chunk 0 at +0, K chunk 1 at +2048 B
  uint4* rowx = &sA[0][(tid >> 3) * 8 + (tid & 7)];
  uint4* rowy = &sA[1][(tid >> 3) * 8 + (tid & 7)];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  const uint64_t dax = umma_desc_kmajor(smem_u32(&sA[0][0]), 128 * 16, 128), day = umma_desc_kmajor(smem_u32(&sA[1][0]), 128 * 16, 128);

  // ---- persistent loop over this CTA's units u = blockIdx.x, +gridDim.x, ...; a "step" is one batch of 128
  // Gaussians of a unit.  The record of the next step (possibly the first batch of the next unit) and the planes
  // of the next unit are fetched one step / one unit ahead.
  struct Unit { int tile, start, n; };
  auto decode = [&](const int4& d) -> Unit { return Unit{d.x, d.y, d.z}; };
  const int4 dzero = make_int4(0, 0, 0, 0);
  auto fetch_planes = [&](int tile, int buf) {             // ONE bulk copy (async proxy), lands on bar_load[buf]
    mbar_expect_tx(&bar_load[buf], PLANE_BYTES);
    bulk_g2s(&sP[buf][0], planes + (size_t)tile * (PLANE_STRIDE / 16), PLANE_BYTES, &bar_load[buf]);
  };
  // Staging as in the forward: the Gaussian id of a step is a coalesced load into a REGISTER two steps ahead, the
  // record of a step a cp.async gather into the thread's own slot one step ahead (slots are private: cp.async.wait_group
  // orders them without a barrier, and no register waits on a scattered global load).  id -1 = slot past the list.
  auto id_of = [&](const Unit& q, int batch) -> int {
    const int i = batch * BT_THREADS + tid;
    return i < q.n ? __ldg(vals + q.start + i) : -1;
  };
  auto fetch_rec = [&](int id) {
    if (id >= 0) {
      const float4* src = rec + 3 * (size_t)id;
      cp_async16_b(&sRec[0][tid], src);                     // two of the record's three float4: the clamped colour
      cp_async16_b(&sRec[1][tid], src + 1);                 // is rebuilt from its fp16 hi | lo halves below
      if (DEPTH) cp_async16_b(&sRec[DEPTH ? 2 : 0][tid], src + 2);   // zabs rides in the third one
    }
    cp_async_commit_b();
  };
  auto nbatch_of = [&](const Unit& q) { return (q.n + BT_THREADS - 1) / BT_THREADS; };
  int u = blockIdx.x;
  // descriptors run three units ahead in registers: a unit costs one 16-byte load whose latency nobody waits for
  Unit cur = decode(__ldg(udesc + u));
  int un = u + gridDim.x, unn = un + gridDim.x;
  int4 d_nxt = un < nunits ? __ldg(udesc + un) : dzero;
  int4 d_nn = unn < nunits ? __ldg(udesc + unn) : dzero;
  if (tid == 0) fetch_planes(cur.tile, 0);
  int id_cur = id_of(cur, 0), id1 = -1;                    // ids of steps 0 and 1, record of step 0
  fetch_rec(id_cur);
  if (1 < nbatch_of(cur)) id1 = id_of(cur, 1);
  else if (un < nunits) id1 = id_of(decode(d_nxt), 0);
  // sums of the previous step waiting for their REDs: Gaussian id (complemented when op == 0: only S is kept), 8 sums
  int pend_id = INT_MIN;
  float pend[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  auto flush_sums = [&]() {
    if (pend_id == INT_MIN) return;
    const bool zop = pend_id < 0;
    float* dst = gacc + (size_t)(zop ? ~pend_id : pend_id) * GACC_F;
    if (!zop) {
      red_add_v4(dst, pend[0], pend[1], pend[2], pend[7]);
      red_add_v4(dst + 4, pend[3], pend[4], pend[5], pend[6]);
      if (DEPTH) atomicAdd(dst + 8, pend[8]);
    } else {
      red_add_v4(dst + 4, pend[3], 0.0f, 0.0f, 0.0f);
    }
    pend_id = INT_MIN;
  };
  uint32_t phase = 0;
  int kbuf = 0;                                            // units processed so far: plane buffer = kbuf & 1
  while (u < nunits) {
    // the next unit of this CTA: its planes and ids are fetched a whole unit ahead
    const Unit nxt = decode(d_nxt), nn = decode(d_nn);
    const int un3 = unn + gridDim.x;
    const int4 d_n3 = un3 < nunits ? __ldg(udesc + un3) : dzero;           // consumed when this unit is done
    const int nb_nxt = nbatch_of(nxt);
    if (un < nunits && tid == 0) fetch_planes(nxt.tile, (kbuf + 1) & 1);   // that buffer's last reader (unit kbuf-1) has retired
    const float k_us = __ldg(tile_scale + cur.tile);      // 2^-(sG+16): planes' 2^sG and the 2^8 of each factor
    const int tx = cur.tile % vp.tiles_x, ty = cur.tile / vp.tiles_x;
    const float x0 = (float)(tx * TILE) + 0.5f, y0 = (float)(ty * TILE) + 0.5f;
    const uint32_t pb = smem_u32(&sP[kbuf & 1][0]);
    const int nbatch = (cur.n + BT_THREADS - 1) / BT_THREADS;
    for (int bi = 0; bi < nbatch; ++bi) {
      cp_async_wait_b<0>();                                // this step's record (issued after the previous step's barrier)
      const int cur_id = id_cur;
      const bool active = cur_id >= 0;
      float4 ra = make_float4(1e18f, -1.0f, 0.0f, 0.0f), rb = ra, col = make_float4(0.f, 0.f, 0.f, 0.f);
      if (active) {
        ra = sRec[0][tid];
        rb = sRec[1][tid];
        // clamped colour = hi + lo of the forward's pre-split halves (22 significant bits): red in ra.w, green in
        // rb.w, blue in rb.z, each {hi | lo << 16} -- saves the gather of the record's third float4
        const uint32_t cr = __float_as_uint(ra.w), cg = __float_as_uint(rb.w), cb = __float_as_uint(rb.z);
        col.x = __half2float(__ushort_as_half((unsigned short)(cr & 0xffffu))) + __half2float(__ushort_as_half((unsigned short)(cr >> 16)));
        col.y = __half2float(__ushort_as_half((unsigned short)(cg & 0xffffu))) + __half2float(__ushort_as_half((unsigned short)(cg >> 16)));
        col.z = __half2float(__ushort_as_half((unsigned short)(cb & 0xffffu))) + __half2float(__ushort_as_half((unsigned short)(cb >> 16)));
        if (DEPTH) col.w = sRec[DEPTH ? 2 : 0][tid].w;
      }
      const float lop = ra.z;
      // ---- factors of this thread's Gaussian, scaled by 2^8 (fp16 range), WITHOUT opacity
      const float dx0 = x0 - ra.x, dy0 = y0 - rb.x;
      float2 fx2[8], fy2[8];
      if (RECUR) {       // by recurrence from the tile centre: 14 MUFU.EX2 + 26 packed multiplies (common.cuh)
        factors16(ra.y, dx0, 8.0f, fx2);
        factors16(rb.y, dy0, 8.0f, fy2);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float dxa = dx0 + (float)(2 * j), dxb = dx0 + (float)(2 * j + 1);
          const float dya = dy0 + (float)(2 * j), dyb = dy0 + (float)(2 * j + 1);
          fx2[j] = make_float2(ex2_approx(fmaf(ra.y * dxa, dxa, 8.0f)), ex2_approx(fmaf(ra.y * dxb, dxb, 8.0f)));
          fy2[j] = make_float2(ex2_approx(fmaf(rb.y * dya, dya, 8.0f)), ex2_approx(fmaf(rb.y * dyb, dyb, 8.0f)));
        }
      }
      rowx[0]   = make_uint4(pack_h2(fx2[0].x, fx2[0].y), pack_h2(fx2[1].x, fx2[1].y), pack_h2(fx2[2].x, fx2[2].y), pack_h2(fx2[3].x, fx2[3].y));
      rowx[128] = make_uint4(pack_h2(fx2[4].x, fx2[4].y), pack_h2(fx2[5].x, fx2[5].y), pack_h2(fx2[6].x, fx2[6].y), pack_h2(fx2[7].x, fx2[7].y));
      rowy[0]   = make_uint4(pack_h2(fy2[0].x, fy2[0].y), pack_h2(fy2[1].x, fy2[1].y), pack_h2(fy2[2].x, fy2[2].y), pack_h2(fy2[3].x, fy2[3].y));
      rowy[128] = make_uint4(pack_h2(fy2[4].x, fy2[4].y), pack_h2(fy2[5].x, fy2[5].y), pack_h2(fy2[6].x, fy2[6].y), pack_h2(fy2[7].x, fy2[7].y));
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy stores -> visible to the tensor core
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      __syncthreads();                                     // every row written; every thread done reading TMEM (previous step)
      if (tid == 0) {
        if (bi == 0) mbar_wait(&bar_load[kbuf & 1], (uint32_t)(kbuf >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        umma_f16(tmem,      dax, umma_desc_kmajor(pb,            NR * 16, 128), IDESC, 0);    // U  = fx . G_hi
        umma_f16(tmem,      dax, umma_desc_kmajor(pb + 2048,     NR * 16, 128), IDESC, 1);    // U += fx . G_lo
        umma_f16(tmem + NR, day, umma_desc_kmajor(pb + 2 * 2048, NR * 16, 128), IDESC, 0);    // V  = fy . G_hi
        umma_f16(tmem + NR, day, umma_desc_kmajor(pb + 3 * 2048, NR * 16, 128), IDESC, 1);    // V += fy . G_lo
        if constexpr (DEPTH) {
          const uint32_t pd = pb + PLANE4_BYTES;
          umma_f16(tmem_d,      dax, umma_desc_kmajor(pd,            16 * 16, 128), IDESC_D, 0);    // U_D  = fx . gD_hi
          umma_f16(tmem_d,      dax, umma_desc_kmajor(pd + 512,      16 * 16, 128), IDESC_D, 1);    // U_D += fx . gD_lo
          umma_f16(tmem_d + 16, day, umma_desc_kmajor(pd + 2 * 512,  16 * 16, 128), IDESC_D, 0);    // V_D  = fy . gD_hi
          umma_f16(tmem_d + 16, day, umma_desc_kmajor(pd + 3 * 512,  16 * 16, 128), IDESC_D, 1);    // V_D += fy . gD_lo
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar_mma)) : "memory");
      }
      // ---- global memory traffic of this thread goes HERE, behind the barrier: fence.proxy.async is a MEMBAR for the
      // issuing thread, i.e. it waits for every global operation the thread still has in flight.  Issued before the
      // fence (as in v7) the record gather of the next step and the REDs of the previous one had to complete inside
      // the step that issued them -- a full L2 round trip on the critical path of every step.  Issued here they have
      // the MMA round trip, the epilogue and the next step's factor evaluation to land.
      flush_sums();
      fetch_rec(id1);                                      // the next step's record, into the slot read at the top
      id_cur = id1;
      // id of the step after next: in this unit, the next one, or (single-step next unit) the one after it
      if (bi + 2 < nbatch) id1 = id_of(cur, bi + 2);
      else if (un >= nunits) id1 = -1;
      else if (bi + 2 - nbatch < nb_nxt) id1 = id_of(nxt, bi + 2 - nbatch);
      else id1 = unn < nunits ? id_of(nn, 0) : -1;
      mbar_wait(&bar_mma, phase);
      phase ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");

      // ---- thread-local epilogue: this Gaussian's 16 rows (U) and 16 columns (V), packed f32x2
      BwdSums A;
      umma_epilogue<DEPTH>(taddr, taddr_d, dx0, dy0, bc2(col.x), bc2(col.y), bc2(col.z), bc2(col.w),
                           [&](int p) { return fy2[p]; }, [&](int p) { return fx2[p]; }, A);
      const float2 aR = A.aR, aG = A.aG, aB = A.aB, aS = A.aS, aSy = A.aSy, aSyy = A.aSyy, aSx = A.aSx, aSxx = A.aSxx;
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");   // TMEM reads ordered before the next step's barrier
      // the sums of this step are added to gacc after the NEXT barrier (see above)
      if (cur_id >= 0) {
        // op == 0 (log2 op = -inf): forward weight 0, but clamp_min(0) passes dL/dop = sum E*t: keep S with op = 1
        const bool zop = (lop == -INFINITY);
        const float opk = zop ? k_us : ex2_approx(lop) * k_us;   // opacity and the 2^-(sG+16) un-scaling, once
        const float z = zop ? 0.0f : opk;
        pend_id = zop ? ~cur_id : cur_id;
        pend[0] = (aR.x + aR.y) * z; pend[1] = (aG.x + aG.y) * z; pend[2] = (aB.x + aB.y) * z;
        pend[3] = (aS.x + aS.y) * opk; pend[4] = (aSx.x + aSx.y) * z; pend[5] = (aSxx.x + aSxx.y) * z;
        pend[6] = (aSy.x + aSy.y) * z; pend[7] = (aSyy.x + aSyy.y) * z;
        if (DEPTH) pend[8] = (A.aZ.x + A.aZ.y) * z;
      }
    }
    u = un;
    un = unn;
    unn = un3;
    cur = nxt;
    d_nxt = d_nn;
    d_nn = d_n3;
    ++kbuf;
  }
  flush_sums();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(128) : "memory");
    if constexpr (DEPTH) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"(32) : "memory");
  }
}

// Per-view statistics of the depth image for the fit script's depth term (python/fit_multiview_stub.py:298-303):
// stats = {M = max_p depth_p, number of pixels attaining M, sum_p sign(depth_p/(M+1e-6) - d_gt_p) depth_p}, with
// depth_p = max(D/(W+1e-6), 0) from the saved accumulators.  Two passes (the second needs M); float max through the
// integer order of non-negative floats.
__global__ void __launch_bounds__(256)
depth_max_kernel(const float* __restrict__ acc, int hw, float* __restrict__ stats) {
  float m = 0.0f;
  for (int p = blockIdx.x * 256 + threadIdx.x; p < hw; p += gridDim.x * 256)
    m = fmaxf(m, fmaxf(acc[4 * (size_t)hw + p] / (acc[3 * (size_t)hw + p] + 1e-6f), 0.0f));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(stats), __float_as_int(m));
}
__global__ void __launch_bounds__(256)
depth_sums_kernel(const float* __restrict__ acc, const float* __restrict__ depth_gt, int hw, float* __restrict__ stats) {
  const float M = stats[0], im = 1.0f / (M + 1e-6f);
  float cnt = 0.0f, ssum = 0.0f;
  for (int p = blockIdx.x * 256 + threadIdx.x; p < hw; p += gridDim.x * 256) {
    const float dep = fmaxf(acc[4 * (size_t)hw + p] / (acc[3 * (size_t)hw + p] + 1e-6f), 0.0f);
    const float dd = dep * im - depth_gt[p];
    ssum += (float)((dd > 0.f) - (dd < 0.f)) * dep;
    cnt += (dep == M) ? 1.0f : 0.0f;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(stats + 1, cnt);
    atomicAdd(stats + 2, ssum);
  }
}

// Clears the per-view backward sums and plants the colour clamp mask of each Gaussian (written by
// preprocess_kernel into the view state) into the spare slot 9 of its row, where the chain-rule kernel finds it
// next to the sums -- the blend kernel itself never touches it.
__global__ void __launch_bounds__(256)
gacc_init_kernel(const uint8_t* __restrict__ cmask, float4* __restrict__ gacc, int n) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float m = __int_as_float((int)cmask[i]);
  gacc[3 * (size_t)i] = make_float4(0.f, 0.f, 0.f, 0.f);
  gacc[3 * (size_t)i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
  gacc[3 * (size_t)i + 2] = make_float4(0.f, m, 0.f, 0.f);
}

// Row layout (12 floats): {dR, dG, dB, Syy | S, Sx, Sxx, Sy | dZ, colour clamp mask, -, -}: the eight sums every
// backward pass produces leave a thread as TWO 16-byte REDs; only a pass that carries a depth gradient adds a third,
// scalar one (dZ).  (With Syy in slot 8 and dZ in slot 3 every pass paid three: 24.1 -> 22.3 ms per C4 iteration.)
int launch_gacc_init(const uint8_t* cmask, float* gacc, int n, cudaStream_t st) {
  if (n <= 0) return B2S_OK;
  gacc_init_kernel<<<(n + 255) / 256, 256, 0, st>>>(cmask, reinterpret_cast<float4*>(gacc), n);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

static bool use_simt_bwd() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("B2S_BWD_SIMT"); v = (e != nullptr && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

int launch_blend_wsum_bwd(const ViewParams& vp, const float4* rec, const int* vals, const int2* ranges,
                          const int* unit_start, const int2* units, const int4* udesc, const Counters* counters,
                          int64_t unit_cap, const float* acc,
                          const float* g_rgb, const float* g_alpha, const float* g_depth, const FitLossArgs* fl,
                          float* gbuf, float* gacc, cudaStream_t st) {
  if (vp.n_tiles <= 0) return B2S_OK;
  const bool fit_depth = fl != nullptr && fl->depth_gt != nullptr && fl->w_depth != 0.0f;
  const bool depth = g_depth != nullptr || fit_depth;
  if (use_simt_bwd() && fl == nullptr) {   // development cross-check: the FP32-pipe kernel (v3)
    count_path(PATH_BWD_OTHER);
    if (depth) gbuf_kernel<true><<<vp.n_tiles, TILE_PIX, 0, st>>>(vp, acc, g_rgb, g_alpha, g_depth, gbuf);
    else       gbuf_kernel<false><<<vp.n_tiles, TILE_PIX, 0, st>>>(vp, acc, g_rgb, g_alpha, g_depth, gbuf);
    B2S_LAUNCH_CHECK();
    if (depth) blend_wsum_bwd_kernel<true><<<(int)unit_cap, BB_THREADS, 0, st>>>(vp, rec, vals, ranges, unit_start, units, gbuf, gacc);
    else       blend_wsum_bwd_kernel<false><<<(int)unit_cap, BB_THREADS, 0, st>>>(vp, rec, vals, ranges, unit_start, units, gbuf, gacc);
  } else {
    // gbuf region: [tile][10 KB] plane fragments / operand matrices, then one scale per tile
    uint32_t* frag = reinterpret_cast<uint32_t*>(gbuf);
    float* tile_scale = gbuf + (size_t)vp.n_tiles * GBUF_FRAG_WORDS;
    // the tcgen05 kernel is the default; B2S_BWD_MMASYNC=1 keeps the mma.sync kernel (development cross-check)
    static const bool mmasync = [] { const char* e = getenv("B2S_BWD_MMASYNC"); return e != nullptr && e[0] == '1'; }();
    const bool umma = !mmasync;
    float* dstats = nullptr;
    if (fit_depth) {
      // per-view depth statistics (max, arg-max count, signed sum) live behind the per-tile scales
      dstats = tile_scale + vp.n_tiles;
      const int hw = vp.width * vp.height;
      int blocks = (hw + 255) / 256;
      if (blocks > sm_count() * 8) blocks = sm_count() * 8;
      B2S_CUDA_TRY(cudaMemsetAsync(dstats, 0, 4 * sizeof(float), st));
      depth_max_kernel<<<blocks, 256, 0, st>>>(acc, hw, dstats);
      B2S_LAUNCH_CHECK();
      depth_sums_kernel<<<blocks, 256, 0, st>>>(acc, fl->depth_gt, hw, dstats);
      B2S_LAUNCH_CHECK();
    }
#define B2S_GBUF(DD, LL, UU)                                                                                              \
  gbuf_frag_kernel<DD, LL, UU><<<vp.n_tiles, TILE_PIX, 0, st>>>(vp, acc, g_rgb, g_alpha, g_depth, fl ? fl->tgt : nullptr,  \
                                                                fl ? fl->mask : nullptr, fit_depth ? fl->depth_gt : nullptr, dstats, \
                                                                fl ? fl->w_sil : 0.f, fit_depth ? fl->w_depth : 0.f,        \
                                                                fl ? fl->scale : 0.f, fl ? fl->loss_accum : nullptr, frag, tile_scale, \
                                                                fl ? fl->tgt_u8 : nullptr, fl ? fl->mask_u8 : nullptr)
    if (fl != nullptr) {
      if (depth) { if (umma) B2S_GBUF(true, true, true); else B2S_GBUF(true, true, false); }
      else       { if (umma) B2S_GBUF(false, true, true); else B2S_GBUF(false, true, false); }
    } else {
      if (depth) { if (umma) B2S_GBUF(true, false, true); else B2S_GBUF(true, false, false); }
      else       { if (umma) B2S_GBUF(false, false, true); else B2S_GBUF(false, false, false); }
    }
#undef B2S_GBUF
    B2S_LAUNCH_CHECK();
    if (umma) {
      // persistent: 4 CTAs per SM (TMEM: 4 x 128 columns; 3 x 160 with the gD plane), each strides over the unit descriptor table
      static const int cps_env = [] { const char* e = getenv("B2S_BWD_CPS"); const int v = e ? atoi(e) : 0; return (v >= 1 && v <= 4) ? v : 0; }();
      const int cps_max = depth ? 3 : 4;
      const int cps = (cps_env >= 1 && cps_env <= cps_max) ? cps_env : cps_max;
      const int grid = (int)(unit_cap < cps * sm_count() ? unit_cap : cps * sm_count());
      count_path(PATH_BWD_UMMA);
      // B2S_BWD_EX2=1: every factor from its own MUFU.EX2 instead of the recurrence (development cross-check)
      static const bool direct = [] { const char* e = getenv("B2S_BWD_EX2"); return e != nullptr && e[0] == '1'; }();
#define B2S_BWU(RR, DD)                                                                                   \
  blend_wsum_bwd_umma_kernel<RR, DD><<<grid, BT_THREADS, 0, st>>>(vp, rec, vals, udesc, counters,          \
                                                                 reinterpret_cast<const uint4*>(frag), tile_scale, gacc)
      // with a depth gradient every factor comes from its own MUFU.EX2 (see launch_blend_wsum_fwd: the recurrence drops
      // far tails that depth = D/(W+1e-6) amplifies)
      static const bool recur_depth = [] { const char* e = getenv("B2S_DEPTH_RECUR"); return e != nullptr && e[0] == '1'; }();
      if (depth) { if (recur_depth) B2S_BWU(true, true); else B2S_BWU(false, true); }
      else       { if (direct) B2S_BWU(false, false); else B2S_BWU(true, false); }
#undef B2S_BWU
      B2S_LAUNCH_CHECK();
      return B2S_OK;
    }
    count_path(PATH_BWD_OTHER);
    const int blocks = (int)((unit_cap + BM_WARPS - 1) / BM_WARPS);
    const uint4* f4 = reinterpret_cast<const uint4*>(frag);
    static const bool minb4 = [] { const char* e = getenv("B2S_BWD_MINB"); return e != nullptr && e[0] == '4'; }();
    if (depth) blend_wsum_bwd_mma_kernel<true, 1><<<blocks, BM_WARPS * 32, 0, st>>>(vp, rec, vals, ranges, unit_start, units, f4, tile_scale, gacc);
    else if (minb4) blend_wsum_bwd_mma_kernel<false, 4><<<blocks, BM_WARPS * 32, 0, st>>>(vp, rec, vals, ranges, unit_start, units, f4, tile_scale, gacc);
    else       blend_wsum_bwd_mma_kernel<false, 3><<<blocks, BM_WARPS * 32, 0, st>>>(vp, rec, vals, ranges, unit_start, units, f4, tile_scale, gacc);
  }
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

}  // namespace b2s
