// Forward blending.  The reference scatters 4 global atomicAdds per (Gaussian,pixel) pair
// (src/renderer.cu:98-102) or runs 10+ elementwise passes over (256,H,W) temporaries
// (python/torch_renderer.py:167-190); here every pixel accumulates in registers.
//
//   WSUM  : A += w c ; W += w ; D += w z ; out = clamp((bg+A)/(1+W))   torch_renderer.py:181-202
//   SORTED: front-to-back "over" with per-pixel alpha state           renderer_cpu.cpp:196-215,241-257
//
// WSUM schedule (v3): the work unit is (tile, segment of <= SEG Gaussians); ONE WARP owns a unit
// and every lane owns 8 pixels (a column of 8 rows).  The unit's records are gathered with
// cp.async into a per-warp 3-stage ring (no block barriers).  The weight of the reference's
// axis-aligned Gaussians is separable, so the 32 lanes compute the tile's 16 column + 16 row
// factors (1 MUFU.EX2 per lane per Gaussian) and each pixel-pair costs 5 FP32 ops issued as
// packed f32x2 instructions.  Because the weighted sum is order independent, a tile whose list
// spans several units is summed from per-unit partial accumulators by finalize_kernel (fixed
// order: deterministic).  Bound: FP32 pipe, ~5.5 lane-cycles per evaluated pixel-pair.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

namespace b2s {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// ---- shared epilogue: accumulators -> outputs (torch_renderer.py:194-202, renderer_cpu.cpp:222-239)
__device__ __forceinline__ void write_pixel(const ViewParams& vp, size_t p, size_t hw, float R, float G, float B,
                                            float W, float D, float* out_rgb, float* out_alpha, float* out_depth,
                                            float* acc, uint8_t* out_rgba) {
  const float inv = 1.0f / (1.0f + W);
  const float o0 = fminf(fmaxf((view_bg(vp, 0) + R) * inv, 0.0f), 1.0f);
  const float o1 = fminf(fmaxf((view_bg(vp, 1) + G) * inv, 0.0f), 1.0f);
  const float o2 = fminf(fmaxf((view_bg(vp, 2) + B) * inv, 0.0f), 1.0f);
  if (out_rgb != nullptr) {
    out_rgb[3 * p] = o0; out_rgb[3 * p + 1] = o1; out_rgb[3 * p + 2] = o2;
  }
  if (out_alpha != nullptr) out_alpha[p] = fminf(fmaxf(W * inv, 0.0f), 1.0f);
  if (out_depth != nullptr) out_depth[p] = fmaxf(D / (W + 1e-6f), 0.0f);
  if (acc != nullptr) {
    acc[p] = R; acc[hw + p] = G; acc[2 * hw + p] = B; acc[3 * hw + p] = W; acc[4 * hw + p] = D;
  }
  if (out_rgba != nullptr) {
    uchar4 u;
    u.x = (unsigned char)(o0 * 255.0f + 0.5f);
    u.y = (unsigned char)(o1 * 255.0f + 0.5f);
    u.z = (unsigned char)(o2 * 255.0f + 0.5f);
    u.w = 255;
    reinterpret_cast<uchar4*>(out_rgba)[p] = u;
  }
}

constexpr int FW_WARPS = 4;     // units per CTA (independent warps)
constexpr int FW_CHUNK = 32;    // Gaussians per stage: one per lane
constexpr int FW_STAGES = 3;
constexpr int FW_ROWS = 8;      // pixels per lane

struct FwStage {
  float4 a[FW_CHUNK];   // x record  {px, qx, lop|op, bbox x}
  float4 b[FW_CHUNK];   // y record  {py, qy, 0|1,    bbox y}
  float4 c[FW_CHUNK];   // {r, g, b, zabs}
};

__device__ __forceinline__ float2 bcast2(float v) { return make_float2(v, v); }

// The weight is separable (axis-aligned Gaussians): w(x,y) = fx(x) * fy(y) with
//   fx = 2^(qx dx^2 + log2 op),  fy = 2^(qy dy^2).
// Per Gaussian the 32 lanes compute the 16 column factors and the 16 row factors of the tile --
// ONE MUFU.EX2 per lane instead of one per pixel -- exchange them through 128 B of shared
// memory, and every pixel-pair then costs FMUL + FADD + 3 FFMA, issued as packed f32x2
// instructions (FFMA2/FADD2/FMUL2: half the issue slots on sm_100).
// EXACT=false : every pixel of the tile is evaluated
// EXACT=true  : fx/fy are zero outside the Gaussian's pixel bbox and w < 1e-5 is dropped
//               (renderer_cpu.cpp:107-113 -- the native weighted-sum mode; records hold op, not log2 op)
template <bool DEPTH, bool EXACT>
__global__ void __launch_bounds__(FW_WARPS * 32)
blend_wsum_fwd_kernel(const ViewParams vp, const float4* __restrict__ rec, const int* __restrict__ vals,
                      const int2* __restrict__ ranges, const int* __restrict__ unit_start,
                      const int2* __restrict__ units, float* __restrict__ partial, float* __restrict__ out_rgb,
                      float* __restrict__ out_alpha, float* __restrict__ out_depth, float* __restrict__ acc,
                      uint8_t* __restrict__ out_rgba) {
  __shared__ __align__(16) FwStage ring[FW_WARPS][FW_STAGES];
  __shared__ __align__(16) float fac[FW_WARPS][2][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u = blockIdx.x * FW_WARPS + warp;
  if (u >= unit_start[vp.n_tiles]) return;           // warps are independent: no block barrier below
  const int2 ud = units[u];
  const int tile = ud.x;
  const int2 rg = ranges[tile];
  const int start = rg.x + ud.y * vp.seg;
  const int n = max(0, min(vp.seg, rg.y - start));
  const int nseg = unit_start[tile + 1] - unit_start[tile];
  const int tx = tile % vp.tiles_x, ty = tile / vp.tiles_x;
  const int cx = lane & 15, half = lane >> 4;
  const int xi = tx * TILE + cx, yi0 = ty * TILE + half * FW_ROWS;
  // lanes 0..15 own column factors, lanes 16..31 own row factors
  const int ci = half ? (ty * TILE + cx) : xi;
  const float coord = ci + 0.5f;
  FwStage* my = ring[warp];

  float2 W2[FW_ROWS / 2], R2[FW_ROWS / 2], G2[FW_ROWS / 2], B2[FW_ROWS / 2], D2[FW_ROWS / 2];
#pragma unroll
  for (int r = 0; r < FW_ROWS / 2; ++r) W2[r] = R2[r] = G2[r] = B2[r] = D2[r] = make_float2(0.f, 0.f);

  const int nchunks = (n + FW_CHUNK - 1) / FW_CHUNK;
  auto issue = [&](int c) {
    if (c < nchunks) {
      const int i = c * FW_CHUNK + lane;
      if (i < n) {
        const int id = __ldg(vals + start + i);
        const float4* src = rec + 3 * (size_t)id;
        FwStage& s = my[c % FW_STAGES];
        cp_async16(&s.a[lane], src);
        cp_async16(&s.b[lane], src + 1);
        cp_async16(&s.c[lane], src + 2);
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int c = 0; c < FW_STAGES - 1; ++c) issue(c);
  for (int c = 0; c < nchunks; ++c) {
    issue(c + FW_STAGES - 1);
    cp_async_wait<FW_STAGES - 1>();
    __syncwarp();
    const FwStage& s = my[c % FW_STAGES];
    const int cnt = min(FW_CHUNK, n - c * FW_CHUNK);
#pragma unroll 2
    for (int j = 0; j < cnt; ++j) {
      // ---- separable factor of this lane (a and b are 512 B apart: one LDS.128, two addresses)
      const float4 h = *(reinterpret_cast<const float4*>(&s.a[j]) + half * FW_CHUNK);
      const float d = coord - h.x;
      float f;
      if constexpr (!EXACT) {
        f = ex2_approx(fmaf(h.y * d, d, half ? 0.0f : h.z));   // the y record's third slot holds packed colour halves
      } else {
        f = h.z * ex2_approx(h.y * d * d);
        const int bb = __float_as_int(h.w);
        f = (ci >= (bb & 0xffff) && ci <= (bb >> 16)) ? f : 0.0f;
      }
      float* fb = fac[warp][j & 1];
      fb[lane] = f;
      __syncwarp();
      const float wx = fb[cx];
      const float4 wya = *reinterpret_cast<const float4*>(fb + 16 + half * FW_ROWS);
      const float4 wyb = *reinterpret_cast<const float4*>(fb + 20 + half * FW_ROWS);
      const float4 col = s.c[j];
      float2 w2[FW_ROWS / 2];
      w2[0] = __fmul2_rn(bcast2(wx), make_float2(wya.x, wya.y));
      w2[1] = __fmul2_rn(bcast2(wx), make_float2(wya.z, wya.w));
      w2[2] = __fmul2_rn(bcast2(wx), make_float2(wyb.x, wyb.y));
      w2[3] = __fmul2_rn(bcast2(wx), make_float2(wyb.z, wyb.w));
#pragma unroll
      for (int r = 0; r < FW_ROWS / 2; ++r) {
        if constexpr (EXACT) {
          w2[r].x = (w2[r].x >= 1e-5f) ? w2[r].x : 0.0f;
          w2[r].y = (w2[r].y >= 1e-5f) ? w2[r].y : 0.0f;
        }
        W2[r] = __fadd2_rn(W2[r], w2[r]);
        R2[r] = __ffma2_rn(w2[r], bcast2(col.x), R2[r]);
        G2[r] = __ffma2_rn(w2[r], bcast2(col.y), G2[r]);
        B2[r] = __ffma2_rn(w2[r], bcast2(col.z), B2[r]);
        if (DEPTH) D2[r] = __ffma2_rn(w2[r], bcast2(col.w), D2[r]);
      }
    }
    __syncwarp();
  }
  cp_async_wait<0>();

  float R[FW_ROWS], G[FW_ROWS], B[FW_ROWS], W[FW_ROWS], D[FW_ROWS];
#pragma unroll
  for (int r = 0; r < FW_ROWS / 2; ++r) {
    R[2 * r] = R2[r].x; R[2 * r + 1] = R2[r].y;
    G[2 * r] = G2[r].x; G[2 * r + 1] = G2[r].y;
    B[2 * r] = B2[r].x; B[2 * r + 1] = B2[r].y;
    W[2 * r] = W2[r].x; W[2 * r + 1] = W2[r].y;
    D[2 * r] = D2[r].x; D[2 * r + 1] = D2[r].y;
  }
  const size_t hw = (size_t)vp.width * vp.height;
  if (nseg <= 1) {
    if (xi < vp.width) {
#pragma unroll
      for (int r = 0; r < FW_ROWS; ++r) {
        const int yi = yi0 + r;
        if (yi < vp.height)
          write_pixel(vp, (size_t)yi * vp.width + xi, hw, R[r], G[r], B[r], W[r], D[r], out_rgb, out_alpha, out_depth,
                      acc, out_rgba);
      }
    }
  } else {
    float* dst = partial + (size_t)u * 5 * TILE_PIX + (half * FW_ROWS) * TILE + cx;
#pragma unroll
    for (int r = 0; r < FW_ROWS; ++r) {
      dst[r * TILE] = R[r];
      dst[TILE_PIX + r * TILE] = G[r];
      dst[2 * TILE_PIX + r * TILE] = B[r];
      dst[3 * TILE_PIX + r * TILE] = W[r];
      if (DEPTH) dst[4 * TILE_PIX + r * TILE] = D[r];
    }
  }
}

// ---- v4 (torch-style weighted sum): the rank-1 updates as GEMMs on the tensor cores ----------------
// With w_i(r,c) = fy_i[r] fx_i[c] every accumulator plane of a tile is a matrix product over the
// tile's Gaussian list (the K dimension):
//   ACC_ch[r][c] = sum_i fy_i[r] * (v_i,ch * fx_i[c]),   v_i = (red, green, blue, 1 [, zabs])
// M = 16 rows, N = 16 columns x CH planes, K = Gaussians: mma.sync m16n8k8, FP32 accumulate.
// The image must match the reference to 1e-4 ABSOLUTE (north_star), which plain TF32 inputs
// (2^-11 relative) do not guarantee, so both operands are split hi + lo (3xTF32):
//   fy fx = fy_hi fx_hi + fy_lo fx_hi + fy_hi fx_lo   (dropped term ~2^-22)
// One warp owns a work unit and consumes 8 Gaussians per step: lane (g = lane/4, t = lane%4)
// evaluates Gaussians t and t+4 of the step at rows g, g+8 and columns g, g+8 (8 MUFU.EX2), which is
// exactly its share of the A (rows x Gaussians) and B (Gaussians x columns) fragments, and ends up
// holding pixels (rows g, g+8) x (columns 2t, 2t+1, 8+2t, 9+2t) of every plane.
// The native "exact" mode (per-pixel w < 1e-5 cut, bbox mask: not separable) stays on the FP32 kernel above.
constexpr int FM_WARPS = 4;
constexpr int FM_STAGE = 32;
constexpr int FM_STAGES = 3;

__device__ __forceinline__ uint32_t tf32_hi(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
__device__ __forceinline__ uint32_t tf32_lo(float x, uint32_t hi) { return __float_as_uint(x - __uint_as_float(hi)) & 0xffffe000u; }
__device__ __forceinline__ void mma_tf32_acc(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <bool DEPTH>
__global__ void __launch_bounds__(FM_WARPS * 32)
blend_wsum_fwd_mma_kernel(const ViewParams vp, const float4* __restrict__ rec, const int* __restrict__ vals,
                          const int2* __restrict__ ranges, const int* __restrict__ unit_start,
                          const int2* __restrict__ units, float* __restrict__ partial, float* __restrict__ out_rgb,
                          float* __restrict__ out_alpha, float* __restrict__ out_depth, float* __restrict__ acc,
                          uint8_t* __restrict__ out_rgba) {
  constexpr int CH = DEPTH ? 5 : 4;      // planes R G B W [D]
  __shared__ __align__(16) FwStage ring[FM_WARPS][FM_STAGES];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u = blockIdx.x * FM_WARPS + warp;
  if (u >= unit_start[vp.n_tiles]) return;             // warps are independent: no block barrier below
  const int2 ud = units[u];
  const int tile = ud.x;
  const int2 rg = ranges[tile];
  const int start = rg.x + ud.y * vp.seg;
  const int n = max(0, min(vp.seg, rg.y - start));
  const int nseg = unit_start[tile + 1] - unit_start[tile];
  const int tx = tile % vp.tiles_x, ty = tile / vp.tiles_x;
  const int g = lane >> 2, t = lane & 3;
  FwStage* my = ring[warp];

  const int nchunks = (n + FM_STAGE - 1) / FM_STAGE;
  auto issue = [&](int c, int id) {
    if (c < nchunks && c * FM_STAGE + lane < n) {
      const float4* src = rec + 3 * (size_t)id;
      FwStage& st = my[c % FM_STAGES];
      cp_async16(&st.a[lane], src);
      cp_async16(&st.b[lane], src + 1);
      cp_async16(&st.c[lane], src + 2);
    }
    cp_async_commit();
  };
  auto load_id = [&](int c) -> int {
    const int i = c * FM_STAGE + lane;
    return (c < nchunks && i < n) ? __ldg(vals + start + i) : 0;
  };
  {
    int ids[FM_STAGES - 1];
#pragma unroll
    for (int c = 0; c < FM_STAGES - 1; ++c) ids[c] = load_id(c);
#pragma unroll
    for (int c = 0; c < FM_STAGES - 1; ++c) issue(c, ids[c]);
  }
  int id_pf = load_id(FM_STAGES - 1);    // ids run one stage ahead of the record gathers

  float D[CH][2][4];
#pragma unroll
  for (int ch = 0; ch < CH; ++ch)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int k = 0; k < 4; ++k) D[ch][h][k] = 0.0f;
  const float cy0 = (float)(ty * TILE + g) + 0.5f, cy1 = cy0 + 8.0f;     // rows g, g+8
  const float cx0 = (float)(tx * TILE + g) + 0.5f, cx1 = cx0 + 8.0f;     // columns g, g+8

  for (int c = 0; c < nchunks; ++c) {
    issue(c + FM_STAGES - 1, id_pf);
    id_pf = load_id(c + FM_STAGES);
    cp_async_wait<FM_STAGES - 1>();
    __syncwarp();
    const FwStage& st = my[c % FM_STAGES];
    const int left = n - c * FM_STAGE;               // Gaussians in this chunk (may exceed 32)
#pragma unroll 2
    for (int sp = 0; sp < FM_STAGE / 8; ++sp) {
      if (sp * 8 >= left) break;                     // warp-uniform
      float fy[2][2], fx[2][2], v[2][4];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = sp * 8 + t + 4 * e;
        const bool on = j < left;                    // padding of the last step: zero weight, clean values
        const float4 ra = st.a[j], rb = st.b[j], rc = st.c[j];
        const float dy0 = cy0 - rb.x, dy1 = cy1 - rb.x, dx0 = cx0 - ra.x, dx1 = cx1 - ra.x;
        const float y0v = ex2_approx(rb.y * dy0 * dy0), y1v = ex2_approx(rb.y * dy1 * dy1);
        const float x0v = ex2_approx(fmaf(ra.y * dx0, dx0, ra.z)), x1v = ex2_approx(fmaf(ra.y * dx1, dx1, ra.z));
        fy[e][0] = on ? y0v : 0.0f; fy[e][1] = on ? y1v : 0.0f;
        fx[e][0] = on ? x0v : 0.0f; fx[e][1] = on ? x1v : 0.0f;
        v[e][0] = on ? rc.x : 0.0f; v[e][1] = on ? rc.y : 0.0f; v[e][2] = on ? rc.z : 0.0f; v[e][3] = on ? rc.w : 0.0f;
      }
      // A = fy: a0 = (row g, Gaussian t), a1 = (row g+8, t), a2 = (row g, t+4), a3 = (row g+8, t+4)
      uint32_t Ah[4], Al[4];
      {
        const float av[4] = {fy[0][0], fy[0][1], fy[1][0], fy[1][1]};
#pragma unroll
        for (int k = 0; k < 4; ++k) { Ah[k] = tf32_hi(av[k]); Al[k] = tf32_lo(av[k], Ah[k]); }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        // B = v * fx at column 8h+g: b0 = Gaussian t, b1 = Gaussian t+4
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) {
          float b0, b1;
          if (ch == 3) { b0 = fx[0][h]; b1 = fx[1][h]; }                       // weight plane
          else if (ch == 4) { b0 = v[0][3] * fx[0][h]; b1 = v[1][3] * fx[1][h]; }   // depth plane
          else { b0 = v[0][ch] * fx[0][h]; b1 = v[1][ch] * fx[1][h]; }
          uint32_t Bh[2], Bl[2];
          Bh[0] = tf32_hi(b0); Bh[1] = tf32_hi(b1);
          Bl[0] = tf32_lo(b0, Bh[0]); Bl[1] = tf32_lo(b1, Bh[1]);
          mma_tf32_acc(D[ch][h], Al, Bh);
          mma_tf32_acc(D[ch][h], Ah, Bl);
          mma_tf32_acc(D[ch][h], Ah, Bh);
        }
      }
    }
    __syncwarp();
  }
  cp_async_wait<0>();

  // accumulator fragment -> pixels: D[ch][h] = {(row g, col 8h+2t), (g, 8h+2t+1), (g+8, 8h+2t), (g+8, 8h+2t+1)}
  const size_t hw = (size_t)vp.width * vp.height;
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = g + 8 * (k >> 1), cc = 8 * h + 2 * t + (k & 1);
      const float R = D[0][h][k], G = D[1][h][k], B = D[2][h][k], W = D[3][h][k], Dz = DEPTH ? D[CH - 1][h][k] : 0.0f;
      if (nseg <= 1) {
        const int xi = tx * TILE + cc, yi = ty * TILE + r;
        if (xi < vp.width && yi < vp.height)
          write_pixel(vp, (size_t)yi * vp.width + xi, hw, R, G, B, W, Dz, out_rgb, out_alpha, out_depth, acc, out_rgba);
      } else {
        float* dst = partial + (size_t)u * 5 * TILE_PIX + r * TILE + cc;
        dst[0] = R;
        dst[TILE_PIX] = G;
        dst[2 * TILE_PIX] = B;
        dst[3 * TILE_PIX] = W;
        if (DEPTH) dst[4 * TILE_PIX] = Dz;
      }
    }
}

// ---- v5: the same GEMM with fp16 operands, K = 16 Gaussians per MMA (3xFP16) ----------------------
// mma.sync m16n8k16 issues at the same rate as the TF32 m16n8k8 (profiles/microbench/mma_rate_b200.txt)
// and consumes twice the Gaussians, so the hi/lo-split product costs 1.5 HMMA per (Gaussian,tile)
// instead of 3.  fp16 has TF32's 11-bit significand; its narrow exponent is handled by scaling both
// factors by 2^8 inside the ex2 argument (free) and un-scaling the accumulators once at the end:
//   fy*2^8 in (0, 256],  v*fx*2^8 <= 256*op*v (clamped to 65000 = op*v up to 253),
//   values below 2^-14 (true factor < 2.4e-7) lose relative precision but stay within 2^-25 absolute.
// Lane (g, t) evaluates Gaussians 2t, 2t+1, 2t+8, 2t+9 of the step at rows g, g+8 and columns g, g+8
// (16 MUFU.EX2): exactly its A (rows x Gaussians) and B (Gaussians x columns) fragment slots.
// Padding slots of the last step hold a record with px = py = 1e18 (factor underflows to 0): no selects.
__device__ __forceinline__ uint32_t h2_bits(__half2 h) { return *reinterpret_cast<const uint32_t*>(&h); }
__device__ __forceinline__ __half2 u32_h2(uint32_t u) { return *reinterpret_cast<const __half2*>(&u); }
__device__ __forceinline__ void split_h2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 f = __half22float2(h);
  hi = h2_bits(h);
  lo = h2_bits(__floats2half2_rn(x0 - f.x, x1 - f.y));
}
__device__ __forceinline__ void mma_f16_acc(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <bool DEPTH>
__global__ void __launch_bounds__(FM_WARPS * 32)
blend_wsum_fwd_f16_kernel(const ViewParams vp, const float4* __restrict__ rec, const int* __restrict__ vals,
                          const int2* __restrict__ ranges, const int* __restrict__ unit_start,
                          const int2* __restrict__ units, float* __restrict__ partial, float* __restrict__ out_rgb,
                          float* __restrict__ out_alpha, float* __restrict__ out_depth, float* __restrict__ acc,
                          uint8_t* __restrict__ out_rgba) {
  constexpr int CH = DEPTH ? 5 : 4;      // planes R G B W [D]
  __shared__ __align__(16) FwStage ring[FM_WARPS][FM_STAGES];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u = blockIdx.x * FM_WARPS + warp;
  if (u >= unit_start[vp.n_tiles]) return;             // warps are independent: no block barrier below
  const int2 ud = units[u];
  const int tile = ud.x;
  const int2 rg = ranges[tile];
  const int start = rg.x + ud.y * vp.seg;
  const int n = max(0, min(vp.seg, rg.y - start));
  const int nseg = unit_start[tile + 1] - unit_start[tile];
  const int tx = tile % vp.tiles_x, ty = tile / vp.tiles_x;
  const int g = lane >> 2, t = lane & 3;
  FwStage* my = ring[warp];

  const int nchunks = (n + FM_STAGE - 1) / FM_STAGE;
  auto issue = [&](int c, int id) {
    if (c < nchunks) {
      FwStage& st = my[c % FM_STAGES];
      if (c * FM_STAGE + lane < n) {
        const float4* src = rec + 3 * (size_t)id;
        cp_async16(&st.a[lane], src);
        cp_async16(&st.b[lane], src + 1);
        cp_async16(&st.c[lane], src + 2);
      } else {   // padding: a record whose factors underflow to exactly 0 (and whose colour halves are 0)
        st.a[lane] = make_float4(1e18f, -1.0f, 0.0f, 0.0f);
        st.b[lane] = make_float4(1e18f, -1.0f, 0.0f, 0.0f);
        st.c[lane] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      }
    }
    cp_async_commit();
  };
  auto load_id = [&](int c) -> int {
    const int i = c * FM_STAGE + lane;
    return (c < nchunks && i < n) ? __ldg(vals + start + i) : 0;
  };
  {
    int ids[FM_STAGES - 1];
#pragma unroll
    for (int c = 0; c < FM_STAGES - 1; ++c) ids[c] = load_id(c);
#pragma unroll
    for (int c = 0; c < FM_STAGES - 1; ++c) issue(c, ids[c]);
  }
  int id_pf = load_id(FM_STAGES - 1);    // ids run one stage ahead of the record gathers

  float D[CH][2][4];
#pragma unroll
  for (int ch = 0; ch < CH; ++ch)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int k = 0; k < 4; ++k) D[ch][h][k] = 0.0f;
  const float cy0 = (float)(ty * TILE + g) + 0.5f, cy1 = cy0 + 8.0f;     // rows g, g+8
  const float cx0 = (float)(tx * TILE + g) + 0.5f, cx1 = cx0 + 8.0f;     // columns g, g+8
  const float2 ncy = make_float2(-cy0, -cy1), ncx = make_float2(-cx0, -cx1);

  for (int c = 0; c < nchunks; ++c) {
    issue(c + FM_STAGES - 1, id_pf);
    id_pf = load_id(c + FM_STAGES);
    cp_async_wait<FM_STAGES - 1>();
    __syncwarp();
    const FwStage& st = my[c % FM_STAGES];
    // both 16-Gaussian steps of the stage run unconditionally (slots past the list hold padding records), so the
    // two steps are one straight-line block the scheduler can interleave: each is a long dependent chain
    // LDS -> exponent -> EX2 -> fp16 split -> HFMA2 -> 3 chained HMMA
#pragma unroll
    for (int sp = 0; sp < FM_STAGE / 16; ++sp) {
      // e = 0..3 -> Gaussians 2t, 2t+1, 2t+8, 2t+9 of the step (the K slots of a0/a1 | a2/a3 and b0 | b1)
      float fy[4][2], fx[4][2], zv[4];
      uint32_t cw[4][3];                             // the clamped colour pre-split as f16 {hi | lo << 16}: r, g, b
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = sp * 16 + 2 * t + (e & 1) + 8 * (e >> 1);
        const float4 ra = st.a[j], rb = st.b[j];
        // exponents of the two rows / two columns as packed f32x2 (the sign of d is irrelevant: it is squared)
        const float2 dy = __fadd2_rn(bcast2(rb.x), ncy), dx = __fadd2_rn(bcast2(ra.x), ncx);
        const float lop8 = fminf(ra.z, 7.99f) + 8.0f;           // fx * 2^8 stays inside fp16 (op <= 253)
        const float2 ay = __ffma2_rn(__fmul2_rn(bcast2(rb.y), dy), dy, bcast2(8.0f));
        const float2 ax = __ffma2_rn(__fmul2_rn(bcast2(ra.y), dx), dx, bcast2(lop8));
        fy[e][0] = ex2_approx(ay.x);
        fy[e][1] = ex2_approx(ay.y);
        fx[e][0] = ex2_approx(ax.x);
        fx[e][1] = ex2_approx(ax.y);
        cw[e][0] = __float_as_uint(ra.w); cw[e][1] = __float_as_uint(rb.w); cw[e][2] = __float_as_uint(rb.z);
        if (DEPTH) zv[e] = st.c[j].w;
      }
      // A = fy: a0 = (row g; K 2t, 2t+1), a1 = (row g+8; same), a2 = (row g; K 2t+8, 2t+9), a3 = (row g+8; same)
      uint32_t Ah[4], Al[4];
      split_h2(fy[0][0], fy[1][0], Ah[0], Al[0]);
      split_h2(fy[0][1], fy[1][1], Ah[1], Al[1]);
      split_h2(fy[2][0], fy[3][0], Ah[2], Al[2]);
      split_h2(fy[2][1], fy[3][1], Ah[3], Al[3]);
      // colour halves paired along K like the B fragment: p = 0 -> Gaussians (2t, 2t+1), p = 1 -> (2t+8, 2t+9)
      __half2 VH[2][3], VL[2][3];
#pragma unroll
      for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          VH[p][ch] = u32_h2(__byte_perm(cw[2 * p][ch], cw[2 * p + 1][ch], 0x5410));
          VL[p][ch] = u32_h2(__byte_perm(cw[2 * p][ch], cw[2 * p + 1][ch], 0x7632));
        }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        // B = v * fx at column 8h+g: b0 = (K 2t, 2t+1), b1 = (K 2t+8, 2t+9).  The weight plane is fx itself,
        // split hi + lo once; a colour plane is formed from the two splits with packed half arithmetic:
        //   hi = rn(vh fxh),  lo = (vh fxh - hi) [exact: one HFMA2] + vh fxl + vl fxh      (vl fxl ~ 2^-22 dropped)
        uint32_t Fh[2], Fl[2];
        split_h2(fx[0][h], fx[1][h], Fh[0], Fl[0]);
        split_h2(fx[2][h], fx[3][h], Fh[1], Fl[1]);
        const __half2 fh0 = u32_h2(Fh[0]), fh1 = u32_h2(Fh[1]), fl0 = u32_h2(Fl[0]), fl1 = u32_h2(Fl[1]);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const __half2 p0 = __hmul2(VH[0][ch], fh0), p1 = __hmul2(VH[1][ch], fh1);
          __half2 l0 = __hfma2(VH[0][ch], fh0, __hneg2(p0)), l1 = __hfma2(VH[1][ch], fh1, __hneg2(p1));
          l0 = __hfma2(VH[0][ch], fl0, l0); l1 = __hfma2(VH[1][ch], fl1, l1);
          l0 = __hfma2(VL[0][ch], fh0, l0); l1 = __hfma2(VL[1][ch], fh1, l1);
          mma_f16_acc(D[ch][h], Al, h2_bits(p0), h2_bits(p1));
          mma_f16_acc(D[ch][h], Ah, h2_bits(l0), h2_bits(l1));
          mma_f16_acc(D[ch][h], Ah, h2_bits(p0), h2_bits(p1));
        }
        mma_f16_acc(D[3][h], Al, Fh[0], Fh[1]);
        mma_f16_acc(D[3][h], Ah, Fl[0], Fl[1]);
        mma_f16_acc(D[3][h], Ah, Fh[0], Fh[1]);
        if (DEPTH) {     // depth plane: zabs is an arbitrary fp32, formed and split in FP32
          uint32_t Bh0, Bl0, Bh1, Bl1;
          split_h2(fminf(zv[0] * fx[0][h], 65000.0f), fminf(zv[1] * fx[1][h], 65000.0f), Bh0, Bl0);
          split_h2(fminf(zv[2] * fx[2][h], 65000.0f), fminf(zv[3] * fx[3][h], 65000.0f), Bh1, Bl1);
          mma_f16_acc(D[CH - 1][h], Al, Bh0, Bh1);
          mma_f16_acc(D[CH - 1][h], Ah, Bl0, Bl1);
          mma_f16_acc(D[CH - 1][h], Ah, Bh0, Bh1);
        }
      }
    }
    __syncwarp();
  }
  cp_async_wait<0>();

  // accumulator fragment -> pixels: D[ch][h] = {(row g, col 8h+2t), (g, 8h+2t+1), (g+8, 8h+2t), (g+8, 8h+2t+1)}
  const size_t hw = (size_t)vp.width * vp.height;
  const float us = 1.0f / 65536.0f;      // the two 2^8 factor scales
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = g + 8 * (k >> 1), cc = 8 * h + 2 * t + (k & 1);
      const float R = D[0][h][k] * us, G = D[1][h][k] * us, B = D[2][h][k] * us, W = D[3][h][k] * us,
                  Dz = DEPTH ? D[CH - 1][h][k] * us : 0.0f;
      if (nseg <= 1) {
        const int xi = tx * TILE + cc, yi = ty * TILE + r;
        if (xi < vp.width && yi < vp.height)
          write_pixel(vp, (size_t)yi * vp.width + xi, hw, R, G, B, W, Dz, out_rgb, out_alpha, out_depth, acc, out_rgba);
      } else {
        float* dst = partial + (size_t)u * 5 * TILE_PIX + r * TILE + cc;
        dst[0] = R;
        dst[TILE_PIX] = G;
        dst[2 * TILE_PIX] = B;
        dst[3 * TILE_PIX] = W;
        if (DEPTH) dst[4 * TILE_PIX] = Dz;
      }
    }
}

// ---- v8: the tile GEMM on the 5th-generation tensor cores (tcgen05 / TMEM) -------------------------------
// Same product as v5, turned so that a THREAD owns a Gaussian (the K index) and the tile's planes are the M x N
// accumulator in tensor memory:
//   D[(plane, column)][row] = sum_i A[(plane, column)][i] * B[row][i],   A = v_plane,i * fx_i[column],  B = fy_i[row]
// 4 planes x 16 columns = 64 operand rows (x hi, lo = M 128), 16 tile rows (x hi, lo = N 32), K = the 128 Gaussians of a batch
// (8 instructions of K = 16).
// Both operands are MN-major in shared memory (thread i stores 16-byte groups of 8 consecutive M / N elements of
// ITS Gaussian: conflict-free STS.128, no transposition), hi/lo split like v5:
//   D[128][32] += [A_hi ; A_lo] . [B_hi | B_lo]      ONE M = 128, N = 32 instruction per 16 Gaussians: hi and lo rows of A
//   stacked along M, hi and lo of B side by side along N; the lo.lo quadrant (~2^-22) is computed and ignored
// The accumulators never pass through registers inside the Gaussian loop: no HMMA issue slots, no fragment
// shuffles, and the per-Gaussian work (32 MUFU.EX2, fp16 splits, 48 colour products) is plain thread-local code.
// Per unit the 128 x 32 accumulator is read back once (lane = operand row), its three useful quadrants summed, exchanged
// through shared memory and written as pixels (or as the unit's partial planes when the tile has several units).
//   CTA = 128 threads, 32 TMEM columns, 44 KB of shared memory -> 5 CTAs per SM; persistent over the units.
// Layout validated by profiles/microbench/umma_probe_mn.cu.  The 5-plane (depth) case stays on v5.
constexpr int FT_THREADS = 128;
constexpr int FT_CTAS = 5;                        // per SM: 44 KB of shared memory, <= 102 registers, 32 TMEM columns each
constexpr uint32_t FT_SBO = FT_THREADS * 16;      // bytes between MN groups of 8: [group][Gaussian] uint4
// DEPTH adds the depth plane D = sum_i z_i fx_i[c] fy_i[r] WITHOUT a fifth operand plane (M = 128 is full): z rides on
// the other operand, B = [fy_hi | fy_lo | (z fy)_hi | (z fy)_lo] (N = 64), and D is read from the weight plane's rows
// (A = fx) at the accumulator columns 32..63.  One instruction per 16 Gaussians as before; +8 KB of shared memory
// (4 CTAs per SM instead of 5) and 64 TMEM columns.
template <bool DEPTH>
struct FtSmemT {
  uint4 Ah[8][FT_THREADS];        // A hi: group = plane * 2 + column / 8
  uint4 Al[8][FT_THREADS];
  uint4 B[DEPTH ? 8 : 4][FT_THREADS];   // fy: hi rows 0-7, hi rows 8-15, lo rows 0-7, lo rows 8-15 [, the same four of z * fy]
  float4 rec[DEPTH ? 3 : 2][FT_THREADS];   // x / y record [/ colour + zabs] of the next step, one private slot per thread
  unsigned long long bar_mma;
  uint32_t tmem_base;
};

// hi + lo split of a packed pair; x - hi is exact in fp32, so the packed FMA (-1 * hi + x) is too
__device__ __forceinline__ void split_h2v(float2 x, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __float22half2_rn(x);
  hi = h2_bits(h);
  lo = h2_bits(__float22half2_rn(__ffma2_rn(__half22float2(h), make_float2(-1.0f, -1.0f), x)));
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
      "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int q = 0; q < 32; ++q) v[q] = __uint_as_float(r[q]);
}

template <bool RECUR, bool DEPTH>
__global__ void __launch_bounds__(FT_THREADS, DEPTH ? FT_CTAS - 1 : FT_CTAS)
blend_wsum_fwd_umma_kernel(const ViewParams vp, const float4* __restrict__ rec, const int* __restrict__ vals,
                           const int2* __restrict__ ranges, const int4* __restrict__ udesc,
                           const Counters* __restrict__ counters, float* __restrict__ partial, float* __restrict__ out_rgb,
                           float* __restrict__ out_alpha, float* __restrict__ out_depth, float* __restrict__ acc,
                           uint8_t* __restrict__ out_rgba) {
  extern __shared__ __align__(128) unsigned char ft_raw[];
  using FtSmem = FtSmemT<DEPTH>;
  FtSmem& sm = *reinterpret_cast<FtSmem*>(ft_raw);
  constexpr int NCOL = DEPTH ? 64 : 32;                                // accumulator columns = operand B rows
  constexpr uint32_t IDESC = umma_idesc_f16(128, NCOL, true, true);    // M = [A_hi ; A_lo] rows, N = [B_hi | B_lo (| zB_hi | zB_lo)] rows
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nunits = counters->n_ne;                       // entries of the unit descriptor table (non-empty units, largest first)
  const size_t hw = (size_t)vp.width * vp.height;
  // ---- tiles without Gaussians are not in the table: their pixels see the background only.  The CTAs share them
  // (tile = blockIdx.x, + gridDim.x, ...; eight range loads in flight at a time).
  for (int t0 = blockIdx.x; t0 < vp.n_tiles; t0 += 8 * gridDim.x) {
    int2 rg[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int t = t0 + e * gridDim.x;
      rg[e] = t < vp.n_tiles ? __ldg(ranges + t) : make_int2(0, 1);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (rg[e].y - rg[e].x > 0) continue;                 // block-uniform
      const int t = t0 + e * gridDim.x;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int pix = tid + h * FT_THREADS;
        const int xi = (t % vp.tiles_x) * TILE + (pix & 15), yi = (t / vp.tiles_x) * TILE + (pix >> 4);
        if (xi < vp.width && yi < vp.height)
          write_pixel(vp, (size_t)yi * vp.width + xi, hw, 0.f, 0.f, 0.f, 0.f, 0.f, out_rgb, out_alpha, out_depth, acc, out_rgba);
      }
    }
  }
  if ((int)blockIdx.x >= nunits) return;                   // block-uniform

  if (tid == 0) mbar_init(&sm.bar_mma, 1);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&sm.tmem_base)), "r"(NCOL) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = sm.tmem_base;
  // K slice s (Gaussians 16 s .. 16 s + 15) starts 256 B into every MN group: K groups 128 B apart (LBO)
  // A_hi and A_lo are adjacent: 16 MN groups of 8 rows, FT_SBO apart = ONE 128-row operand
  const uint64_t dA = umma_desc(smem_u32(&sm.Ah[0][0]), 128, FT_SBO), dB = umma_desc(smem_u32(&sm.B[0][0]), 128, FT_SBO);
  float* sOut = reinterpret_cast<float*>(&sm.Ah[0][0]);    // [plane][row * 16 + column], aliases A hi between units

  struct Unit { int tile, start, n, uidx; };               // uidx: unit index (its partial planes) | tile has several units << 31
  auto decode = [&](const int4& d) -> Unit { return Unit{d.x, d.y, d.z, d.w}; };
  const int4 dzero = make_int4(0, 0, 0, 0);
  // pixels q = tid, tid + 128 of a unit: outputs (single-unit tile) or the unit's partial planes
  auto emit = [&](const Unit& q, int pix, float R, float G, float Bc, float W, float Dz) {
    if (q.uidx >= 0) {
      const int xi = (q.tile % vp.tiles_x) * TILE + (pix & 15), yi = (q.tile / vp.tiles_x) * TILE + (pix >> 4);
      if (xi < vp.width && yi < vp.height)
        write_pixel(vp, (size_t)yi * vp.width + xi, hw, R, G, Bc, W, Dz, out_rgb, out_alpha, out_depth, acc, out_rgba);
    } else {
      float* dst = partial + (size_t)(q.uidx & 0x7fffffff) * 5 * TILE_PIX + pix;
      dst[0] = R;
      dst[TILE_PIX] = G;
      dst[2 * TILE_PIX] = Bc;
      dst[3 * TILE_PIX] = W;
      if (DEPTH) dst[4 * TILE_PIX] = Dz;
    }
  };
  // Staging.  The Gaussian id of a step is a coalesced load into a REGISTER, requested two steps ahead (by the time
  // the next fence -- a MEMBAR for the issuing thread -- comes, it has long landed); the record of a step is
  // gathered with cp.async into the thread's own shared-memory slot one step ahead (the thread has just moved the
  // current record from that slot into registers; slots are private, so cp.async.wait_group orders them without a
  // barrier, and no register waits on a scattered global load).  Nothing else is staged: 44 KB of shared memory
  // per CTA, five CTAs per SM.  id -1 = slot past the unit's list: the padding record.
  auto id_of = [&](const Unit& q, int batch) -> int {
    const int i = batch * FT_THREADS + tid;
    return i < q.n ? __ldg(vals + q.start + i) : -1;
  };
  auto fetch_rec = [&](int id) {
    if (id >= 0) {
      const float4* src = rec + 3 * (size_t)id;
      cp_async16_b(&sm.rec[0][tid], src);
      cp_async16_b(&sm.rec[1][tid], src + 1);
      if (DEPTH) cp_async16_b(&sm.rec[DEPTH ? 2 : 0][tid], src + 2);     // {r, g, b, zabs}: the depth plane needs zabs
    }
    cp_async_commit_b();
  };
  auto nbatch_of = [&](const Unit& q) { return (q.n + FT_THREADS - 1) / FT_THREADS; };

  // descriptors run three units ahead in registers: a unit costs one 16-byte load whose latency nobody waits for
  int u = blockIdx.x;
  Unit cur = decode(__ldg(udesc + u));
  int un = u + gridDim.x, unn = un + gridDim.x;
  int4 d_nxt = un < nunits ? __ldg(udesc + un) : dzero;
  int4 d_nn = unn < nunits ? __ldg(udesc + unn) : dzero;
  // ids of steps 0 and 1, record of step 0
  bool act = false;
  int id1 = -1;
  {
    const int id0 = id_of(cur, 0);
    fetch_rec(id0);
    act = id0 >= 0;
    if (1 < nbatch_of(cur)) id1 = id_of(cur, 1);
    else if (un < nunits) id1 = id_of(decode(d_nxt), 0);
  }
  uint32_t phase = 0;
  bool pending = false;                                    // an MMA batch has been committed and not yet waited for
  while (u < nunits) {
    const Unit nxt = decode(d_nxt), nn = decode(d_nn);
    const int un3 = unn + gridDim.x;
    const int4 d_n3 = un3 < nunits ? __ldg(udesc + un3) : dzero;           // consumed when this unit is done
    const int nb_nxt = nbatch_of(nxt);
    const int tx = cur.tile % vp.tiles_x, ty = cur.tile / vp.tiles_x;
    const float x0 = (float)(tx * TILE) + 0.5f, y0 = (float)(ty * TILE) + 0.5f;
    const int nbatch = (cur.n + FT_THREADS - 1) / FT_THREADS;
    for (int bi = 0; bi < nbatch; ++bi) {
      cp_async_wait_b<0>();                                // this step's record (gathered during the previous step)
      // padding: q = 0 and log2 op = -1000 make every x factor exactly 0 (and every ratio of the recurrence 1: no
      // 0 * inf), the colour halves are 0; a padded K slot then adds 0 to every accumulator
      float4 ra = make_float4(0.0f, 0.0f, -1000.0f, 0.0f), rb = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      float zabs = 0.0f;
      if (act) {
        ra = sm.rec[0][tid];
        rb = sm.rec[1][tid];
        if (DEPTH) zabs = sm.rec[DEPTH ? 2 : 0][tid].w;
      }
      fetch_rec(id1);                                      // next step's record into the slot just read
      act = id1 >= 0;
      // id of the step after next: in this unit, the next one, or (single-step next unit) the one after it
      if (bi + 2 < nbatch) id1 = id_of(cur, bi + 2);
      else if (un >= nunits) id1 = -1;
      else if (bi + 2 - nbatch < nb_nxt) id1 = id_of(nxt, bi + 2 - nbatch);
      else id1 = unn < nunits ? id_of(nn, 0) : -1;
      // ---- factors of this thread's Gaussian at the tile's 16 columns / rows, each scaled by 2^8 (fp16 range),
      // split into fp16 hi + lo pairs {even, odd}
      const float dx0 = ra.x - x0, dy0 = rb.x - y0;
      const float lop8 = fminf(ra.z, 7.99f) + 8.0f;        // fx * 2^8 stays inside fp16 (op <= 253)
      uint32_t Fh[8], Fl[8], Yh[8], Yl[8];
      uint32_t Zh[DEPTH ? 8 : 1], Zl[DEPTH ? 8 : 1];      // z * fy (the depth plane's B rows); fy <= 2^8, so z up to 253 fits fp16
      const float2 z2 = bcast2(zabs), zcap = bcast2(65000.0f);
      auto split_z = [&](int j, float2 fyv) {
        if constexpr (DEPTH) {
          const float2 t = __fmul2_rn(z2, fyv);
          split_h2v(make_float2(fminf(t.x, zcap.x), fminf(t.y, zcap.y)), Zh[j], Zl[j]);
        }
      };
      if (RECUR) {       // by recurrence from the tile centre: 14 MUFU.EX2 + 26 packed multiplies (common.cuh)
        float2 fx2[8], fy2[8];
        factors16(ra.y, -dx0, lop8, fx2);
        factors16(rb.y, -dy0, 8.0f, fy2);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          split_h2v(fx2[j], Fh[j], Fl[j]);
          split_h2v(fy2[j], Yh[j], Yl[j]);
          split_z(j, fy2[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 off = make_float2(-(float)(2 * j), -(float)(2 * j + 1));
          const float2 dx = __fadd2_rn(bcast2(dx0), off), dy = __fadd2_rn(bcast2(dy0), off);
          const float2 ax = __ffma2_rn(__fmul2_rn(bcast2(ra.y), dx), dx, bcast2(lop8));
          const float2 ay = __ffma2_rn(__fmul2_rn(bcast2(rb.y), dy), dy, bcast2(8.0f));
          const float2 fyv = make_float2(ex2_approx(ay.x), ex2_approx(ay.y));
          split_h2v(make_float2(ex2_approx(ax.x), ex2_approx(ax.y)), Fh[j], Fl[j]);
          split_h2v(fyv, Yh[j], Yl[j]);
          split_z(j, fyv);
        }
      }
      // the previous batch's MMAs read the operand buffers: they must have retired before the stores below
      // (they were issued a whole factor computation ago)
      if (pending) { mbar_wait(&sm.bar_mma, phase); phase ^= 1u; pending = false; }
      sm.B[0][tid] = make_uint4(Yh[0], Yh[1], Yh[2], Yh[3]);
      sm.B[1][tid] = make_uint4(Yh[4], Yh[5], Yh[6], Yh[7]);
      sm.B[2][tid] = make_uint4(Yl[0], Yl[1], Yl[2], Yl[3]);
      sm.B[3][tid] = make_uint4(Yl[4], Yl[5], Yl[6], Yl[7]);
      if constexpr (DEPTH) {
        sm.B[4][tid] = make_uint4(Zh[0], Zh[1], Zh[2], Zh[3]);
        sm.B[5][tid] = make_uint4(Zh[4], Zh[5], Zh[6], Zh[7]);
        sm.B[6][tid] = make_uint4(Zl[0], Zl[1], Zl[2], Zl[3]);
        sm.B[7][tid] = make_uint4(Zl[4], Zl[5], Zl[6], Zl[7]);
      }
      sm.Ah[6][tid] = make_uint4(Fh[0], Fh[1], Fh[2], Fh[3]);      // weight plane: fx itself
      sm.Ah[7][tid] = make_uint4(Fh[4], Fh[5], Fh[6], Fh[7]);
      sm.Al[6][tid] = make_uint4(Fl[0], Fl[1], Fl[2], Fl[3]);
      sm.Al[7][tid] = make_uint4(Fl[4], Fl[5], Fl[6], Fl[7]);
      // colour planes from the pre-split clamped colour (record: f16 {hi | lo << 16}):
      //   hi = rn(vh fxh),  lo = (vh fxh - hi) [exact: one HFMA2] + vh fxl + vl fxh      (vl fxl ~ 2^-22 dropped)
      const uint32_t cw[3] = {__float_as_uint(ra.w), __float_as_uint(rb.w), __float_as_uint(rb.z)};
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const __half2 vh = u32_h2(__byte_perm(cw[ch], cw[ch], 0x1010)), vl = u32_h2(__byte_perm(cw[ch], cw[ch], 0x3232));
        uint32_t P[8], L[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const __half2 fh = u32_h2(Fh[j]), fl = u32_h2(Fl[j]);
          const __half2 p = __hmul2(vh, fh);
          __half2 l = __hfma2(vh, fh, __hneg2(p));
          l = __hfma2(vh, fl, l);
          l = __hfma2(vl, fh, l);
          P[j] = h2_bits(p);
          L[j] = h2_bits(l);
        }
        sm.Ah[2 * ch][tid] = make_uint4(P[0], P[1], P[2], P[3]);
        sm.Ah[2 * ch + 1][tid] = make_uint4(P[4], P[5], P[6], P[7]);
        sm.Al[2 * ch][tid] = make_uint4(L[0], L[1], L[2], L[3]);
        sm.Al[2 * ch + 1][tid] = make_uint4(L[4], L[5], L[6], L[7]);
      }
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy stores -> visible to the tensor core
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      __syncthreads();                                     // every operand column written (and, at bi == 0, the previous unit's accumulator read)
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
        for (int s = 0; s < FT_THREADS / 16; ++s) {
          const uint64_t off = (uint64_t)(s * 16);                                  // 256 B in descriptor units
          umma_f16(tmem, dA + off, dB + off, IDESC, (bi > 0 || s > 0) ? 1u : 0u);   // D (+)= [A_hi ; A_lo] . [B_hi | B_lo]
        }
        umma_commit(&sm.bar_mma);
      }
      pending = true;
    }
    // ---- unit epilogue: accumulator -> pixels
    mbar_wait(&sm.bar_mma, phase);                         // nbatch >= 1: the unit is not empty
    phase ^= 1u;
    pending = false;
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    {
      // TMEM lane = operand row: warps 0, 1 hold the hi rows of planes {0,1} / {2,3} (lane = plane % 2 * 16 + column),
      // warps 2, 3 the lo rows of the same planes; TMEM column = row (x B_hi) | 16 + row (x B_lo).  The lo.lo
      // quadrant (warps 2, 3, columns 16..31) is the dropped ~2^-22 term.
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
      float* dst = sOut + ((warp & 1) * 2 + (lane >> 4)) * TILE_PIX + (lane & 15);
      if (warp >= 2) {
#pragma unroll
        for (int r = 0; r < 16; ++r) dst[r * TILE] = v[r];
      }
      // depth plane: the weight plane's rows (operand rows 48..63 hi = warp 1, 112..127 lo = warp 3; lanes 16..31) at the
      // accumulator columns 32..63 = (z fy)_hi | (z fy)_lo rows
      float vz[DEPTH ? 32 : 1];
      float* dstz = sOut + 4 * TILE_PIX + (lane & 15);
      if constexpr (DEPTH) {
        if (warp & 1) {                                    // warp uniform: the collective load runs on whole warps
          tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + 32u, vz);
          if (warp == 3 && lane >= 16) {
#pragma unroll
            for (int r = 0; r < 16; ++r) dstz[r * TILE] = vz[r];
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");   // TMEM reads ordered before the next unit's first MMA
      __syncthreads();
      const float us = 1.0f / 65536.0f;                    // the two 2^8 factor scales
      if (warp < 2) {
#pragma unroll
        for (int r = 0; r < 16; ++r) dst[r * TILE] = (dst[r * TILE] + v[r] + v[16 + r]) * us;
      }
      if constexpr (DEPTH) {
        if (warp == 1 && lane >= 16) {
#pragma unroll
          for (int r = 0; r < 16; ++r) dstz[r * TILE] = (dstz[r * TILE] + vz[r] + vz[16 + r]) * us;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int pix = tid + h * FT_THREADS;
      emit(cur, pix, sOut[pix], sOut[TILE_PIX + pix], sOut[2 * TILE_PIX + pix], sOut[3 * TILE_PIX + pix],
           DEPTH ? sOut[4 * TILE_PIX + pix] : 0.0f);
    }
    __syncthreads();                                       // sOut aliases the A operand: reads done before the next stores
    u = un;
    un = unn;
    unn = un3;
    cur = nxt;
    d_nxt = d_nn;
    d_nn = d_n3;
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(NCOL) : "memory");
}

// Sums the per-unit partial accumulators of tiles that span several units, in unit order.
template <bool DEPTH>
__global__ void __launch_bounds__(TILE_PIX)
finalize_kernel(const ViewParams vp, const int* __restrict__ unit_start, const float* __restrict__ partial,
                float* __restrict__ out_rgb, float* __restrict__ out_alpha, float* __restrict__ out_depth,
                float* __restrict__ acc, uint8_t* __restrict__ out_rgba) {
  const int tile = blockIdx.x;
  const int u0 = unit_start[tile], u1 = unit_start[tile + 1];
  if (u1 - u0 <= 1) return;
  const int q = threadIdx.x;
  float R = 0.f, G = 0.f, B = 0.f, W = 0.f, D = 0.f;
  for (int u = u0; u < u1; ++u) {
    const float* src = partial + (size_t)u * 5 * TILE_PIX + q;
    R += src[0];
    G += src[TILE_PIX];
    B += src[2 * TILE_PIX];
    W += src[3 * TILE_PIX];
    if (DEPTH) D += src[4 * TILE_PIX];
  }
  const int xi = (tile % vp.tiles_x) * TILE + (q & 15), yi = (tile / vp.tiles_x) * TILE + (q >> 4);
  if (xi < vp.width && yi < vp.height)
    write_pixel(vp, (size_t)yi * vp.width + xi, (size_t)vp.width * vp.height, R, G, B, W, D, out_rgb, out_alpha,
                out_depth, acc, out_rgba);
}

int launch_blend_wsum_fwd(const ViewParams& vp, const float4* rec, const int* vals, const int2* ranges,
                          const int* unit_start, const int2* units, const int4* udesc, const Counters* counters,
                          int64_t unit_cap, float* partial,
                          float* out_rgb, float* out_alpha, float* out_depth, float* acc, uint8_t* out_rgba,
                          cudaStream_t st) {
  if (vp.n_tiles <= 0) return B2S_OK;
  const int blocks = (int)((unit_cap + FW_WARPS - 1) / FW_WARPS);
  const bool depth = out_depth != nullptr || vp.keep_depth != 0;   // D is accumulated when the depth image (or its gradient) is wanted
#define B2S_FW(DD, EE)                                                                                              \
  blend_wsum_fwd_kernel<DD, EE><<<blocks, FW_WARPS * 32, 0, st>>>(vp, rec, vals, ranges, unit_start, units, partial, \
                                                                   out_rgb, out_alpha, out_depth, acc, out_rgba)
#define B2S_FWM(KERN, DD)                                                                                             \
  KERN<DD><<<(int)((unit_cap + FM_WARPS - 1) / FM_WARPS), FM_WARPS * 32, 0, st>>>(                                    \
      vp, rec, vals, ranges, unit_start, units, partial, out_rgb, out_alpha, out_depth, acc, out_rgba)
  static const bool simt = [] { const char* e = getenv("B2S_FWD_SIMT"); return e != nullptr && e[0] == '1'; }();
  static const bool mmasync = [] { const char* e = getenv("B2S_FWD_MMASYNC"); return e != nullptr && e[0] == '1'; }();
  static const bool tf32 = [] { const char* e = getenv("B2S_FWD_TF32"); return e != nullptr && e[0] == '1'; }();
  if (vp.exact_bbox) { count_path(PATH_FWD_OTHER); if (depth) B2S_FW(true, true); else B2S_FW(false, true); }
  else if (simt)     { count_path(PATH_FWD_OTHER); if (depth) B2S_FW(true, false); else B2S_FW(false, false); }   // development cross-check (v3)
  else if (!tf32 && !mmasync) {
    // tcgen05: persistent, 5 CTAs per SM (4 with the depth plane), each strides over the unit descriptor table
    B2S_CUDA_TRY(per_device_once(ONCE_FWD_UMMA, [] {
      cudaError_t e = cudaSuccess;
      auto set = [&](auto kern, size_t bytes) { if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes); };
      set(blend_wsum_fwd_umma_kernel<true, false>, sizeof(FtSmemT<false>));
      set(blend_wsum_fwd_umma_kernel<false, false>, sizeof(FtSmemT<false>));
      set(blend_wsum_fwd_umma_kernel<true, true>, sizeof(FtSmemT<true>));
      set(blend_wsum_fwd_umma_kernel<false, true>, sizeof(FtSmemT<true>));
      return e;
    }));
    // B2S_FWD_EX2=1: every factor from its own MUFU.EX2 instead of the recurrence (development cross-check)
    static const bool direct = [] { const char* e = getenv("B2S_FWD_EX2"); return e != nullptr && e[0] == '1'; }();
    static const int cps_env = [] { const char* e = getenv("B2S_FWD_CPS"); const int v = e ? atoi(e) : 0; return (v >= 1 && v <= FT_CTAS) ? v : 0; }();
    const int cps_max = depth ? FT_CTAS - 1 : FT_CTAS;
    const int cps = (cps_env >= 1 && cps_env <= cps_max) ? cps_env : cps_max;
    const int grid = (int)(unit_cap < cps * sm_count() ? unit_cap : cps * sm_count());
    count_path(PATH_FWD_UMMA);
#define B2S_FWU(RR, DD)                                                                                                      \
  blend_wsum_fwd_umma_kernel<RR, DD><<<grid, FT_THREADS, sizeof(FtSmemT<DD>), st>>>(vp, rec, vals, ranges, udesc, counters, partial, \
                                                                                   out_rgb, out_alpha, out_depth, acc, out_rgba)
    // The depth plane always takes the direct evaluation: the recurrence zeroes an axis whose MIDDLE value underflows
    // (Gaussian > 13.6 sigma from the tile centre), which for sigma = 1 px drops weights up to exp(-18) at the tile's
    // near columns -- invisible in rgb / alpha, but depth = D/(W+1e-6) and its gradient amplify exactly those tails
    // (measured 5e-3..1e-2 on the depth-gradient goldens with the recurrence, < 3e-4 without).
    static const bool recur_depth = [] { const char* e = getenv("B2S_DEPTH_RECUR"); return e != nullptr && e[0] == '1'; }();
    if (depth) { if (recur_depth) B2S_FWU(true, true); else B2S_FWU(false, true); }
    else       { if (direct) B2S_FWU(false, false); else B2S_FWU(true, false); }
#undef B2S_FWU
  }
  else if (tf32)     { count_path(PATH_FWD_OTHER); if (depth) B2S_FWM(blend_wsum_fwd_mma_kernel, true); else B2S_FWM(blend_wsum_fwd_mma_kernel, false); }
  else               { count_path(PATH_FWD_OTHER); if (depth) B2S_FWM(blend_wsum_fwd_f16_kernel, true); else B2S_FWM(blend_wsum_fwd_f16_kernel, false); }
#undef B2S_FWM
#undef B2S_FW
  B2S_LAUNCH_CHECK();
  if (depth)
    finalize_kernel<true><<<vp.n_tiles, TILE_PIX, 0, st>>>(vp, unit_start, partial, out_rgb, out_alpha, out_depth, acc, out_rgba);
  else
    finalize_kernel<false><<<vp.n_tiles, TILE_PIX, 0, st>>>(vp, unit_start, partial, out_rgb, out_alpha, out_depth, acc, out_rgba);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

// ---- depth-sorted "over" compositing -----------------------------------------------------------
// (reference src/renderer_cpu.cpp:196-215: a = clamp01(op exp(e)), skip a < 1e-5, contrib = (1 - A) a, C += contrib c,
//  A += contrib, inside the Gaussian's 3-sigma pixel bbox)
// Work unit = (tile, segment of <= SEG consecutive Gaussians of its depth-sorted list), one CTA of 256 threads =
// one pixel per thread.  "Over" is associative: a segment composited alone from A = 0 yields (C_s, A_s), and
//   C = C_1 + (1 - A_1) C_2,  A = A_1 + (1 - A_1) A_2
// so long lists no longer serialise on one CTA (a 960x540 frame of 1 M Gaussians has tiles of > 8000: they were the
// tail of the old one-CTA-per-tile kernel); finalize_sorted_kernel folds a tile's units in order.  Inside a unit the
// CTA stops once every pixel is saturated (1 - A < 1e-4 cannot move an 8-bit channel).
// The weight is separable, a(r,c) = [op 2^(qx dx^2)] [2^(qy dy^2)]: per chunk of 64 Gaussians the 256 threads first
// evaluate the 64 x (16 + 16) factors (8 MUFU.EX2 each, bbox folded in as zeros) into shared memory, then a
// pixel-pair costs 2 LDS + FMUL + compare + ~6 FP32 instead of the exponent arithmetic + MUFU + 4 bbox compares.
constexpr int BS_THREADS = 256;
constexpr int BS_CHUNK = 64;

struct SortedStage {
  float4 a[BS_CHUNK];   // {px, qx, op, bbox x (min | max << 16)}
  float4 b[BS_CHUNK];   // {py, qy, 1, bbox y}
  float4 c[BS_CHUNK];   // {r, g, b, zabs}
};

__device__ __forceinline__ void sorted_stage_chunk(SortedStage& sb, const float4* __restrict__ rec, const int* __restrict__ vals,
                                                   int start, int n, int chunk) {
  const int t = threadIdx.x;
  if (t < BS_CHUNK) {
    const int i = chunk * BS_CHUNK + t;
    if (i < n) {
      const int id = __ldg(vals + start + i);
      const float4* src = rec + 3 * (size_t)id;
      cp_async16(&sb.a[t], src);
      cp_async16(&sb.b[t], src + 1);
      cp_async16(&sb.c[t], src + 2);
    }
  }
}

constexpr int BS_FIRST = 2;      // segments the first kernel walks in order, with early exit, before anything is split

// Two launches.  FIRST: one CTA per tile walks the front BS_FIRST segments of the list in order and stops as soon as
// every pixel is saturated -- most covered tiles end here after a few hundred Gaussians (the lists hold thousands), and
// the tile is marked hidden-behind.  REST: the remaining segments of the tiles that did NOT saturate (silhouette tiles,
// thin coverage), one CTA per segment in parallel; a segment behind a saturating one is skipped (finalize stops folding
// before it).  Splitting everything from the start would composite the hidden 90 % of the interior lists; walking
// everything in order left the frame waiting for the longest non-saturating tile.
template <bool FIRST>
__global__ void __launch_bounds__(BS_THREADS)
blend_sorted_units_kernel(const ViewParams vp, const float4* __restrict__ rec, const int* __restrict__ vals,
                          const int2* __restrict__ ranges, const int* __restrict__ unit_start,
                          float* __restrict__ partial, int* __restrict__ sat_seg, int2* __restrict__ rest_list,
                          float* __restrict__ out_rgb, float* __restrict__ out_alpha, uint8_t* __restrict__ out_rgba,
                          Counters* dbg) {
  __shared__ __align__(16) SortedStage sb[2];
  __shared__ float sfx[BS_CHUNK][TILE], sfy[BS_CHUNK][TILE];
  __shared__ int s_flag;
  // factor phase ownership: thread t evaluates 8 factors of Gaussian t / 4: part 0, 1 = columns 0-7, 8-15; 2, 3 = rows
  const int fj = threadIdx.x >> 2, fpart = threadIdx.x & 3;
  const bool faxis_y = fpart >= 2;
  const int f0 = (fpart & 1) * 8;
  const int col = threadIdx.x & 15, row = threadIdx.x >> 4;
  // Work items: FIRST -- one CTA per tile.  REST -- the (tile, segment) pairs the FIRST kernel found still visible,
  // appended to rest_list ([0].x = count); a persistent grid strides over that list, so no CTA is spent on the empty
  // and the saturated tiles (a CTA per (tile, segment) of the whole frame cost more in launches than in work: 32 640
  // CTAs for 56 live segments in the 960x540 viewer frame).
  const int n_items = FIRST ? 1 : rest_list[0].x;
  for (int w = FIRST ? 0 : (int)blockIdx.x; w < n_items; w += FIRST ? 1 : (int)gridDim.x) {     // block uniform
  const int2 item = FIRST ? make_int2((int)blockIdx.x, 0) : rest_list[1 + w];
  const int tile = item.x, ybase = item.y;
  const int u0 = unit_start[tile];
  const int nseg = unit_start[tile + 1] - u0;
  const int2 rg = ranges[tile];
  const int L = rg.y - rg.x;
  const int tx = tile % vp.tiles_x, ty = tile / vp.tiles_x;
  const int xi = tx * TILE + col, yi = ty * TILE + row;
  const bool inside = xi < vp.width && yi < vp.height;
  const int cbase = (faxis_y ? ty : tx) * TILE + f0;

  bool hidden = false;
  for (int seg = FIRST ? 0 : ybase; seg < (FIRST ? 1 : ybase + 1) && !hidden; ++seg) {   // block uniform: one segment
    if (!FIRST) {
      // a segment in front of this one that saturates every pixel on its own makes it invisible: finalize stops
      // folding before it (the FIRST kernel's verdict is final; among REST segments the race is benign)
      __syncthreads();                                   // everyone has read the previous value of s_flag
      if (threadIdx.x == 0) s_flag = *(volatile int*)(sat_seg + tile);
      __syncthreads();                                   // one read for the whole block: the decision must be uniform
      if (seg > s_flag) { hidden = true; continue; }
    }
    const int e0 = seg * vp.seg;
    const int start = rg.x + e0;
    const int n = max(0, min(FIRST ? BS_FIRST * vp.seg : vp.seg, L - e0));
    const int nchunks = (n + BS_CHUNK - 1) / BS_CHUNK;
    float C0 = 0.f, C1 = 0.f, C2 = 0.f, A = 0.f;
    int saturated = 0;
    __syncthreads();                                     // previous segment's reads of the stage buffers are over
    if (nchunks > 0) sorted_stage_chunk(sb[0], rec, vals, start, n, 0);
    cp_async_commit();
    for (int c = 0; c < nchunks; ++c) {
      if (c + 1 < nchunks) sorted_stage_chunk(sb[(c + 1) & 1], rec, vals, start, n, c + 1);
      cp_async_commit();
      cp_async_wait<1>();
      __syncthreads();                                   // chunk c staged; every thread has left chunk c-1's blend loop
      const SortedStage& s = sb[c & 1];
      const int cnt = min(BS_CHUNK, n - c * BS_CHUNK);
      if (fj < cnt) {
        const float4 h = faxis_y ? s.b[fj] : s.a[fj];
        const int bb = __float_as_int(h.w), lo = bb & 0xffff, hi = bb >> 16;
        float* dst = (faxis_y ? &sfy[fj][0] : &sfx[fj][0]) + f0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int ci = cbase + k;
          const float d = ((float)ci + 0.5f) - h.x;
          const float f = h.z * ex2_approx(h.y * d * d);   // h.z = op on the x record, 1 on the y record
          dst[k] = (ci >= lo && ci <= hi) ? f : 0.0f;
        }
      }
      __syncthreads();
      // a warp whose 32 pixels are all saturated (or outside the image) sits the chunk out
      if (!__all_sync(0xffffffffu, !inside || (1.0f - A) < 5e-5f)) {     // the CTA's own stopping bar, not a looser one
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) {
          float al = sfx[j][col] * sfy[j][row];
          if (al >= 1e-5f) {
            al = fminf(al, 1.0f);
            const float contrib = (1.0f - A) * al;
            if (contrib > 0.0f) {
              const float4 cc = s.c[j];
              C0 = fmaf(contrib, cc.x, C0);
              C1 = fmaf(contrib, cc.y, C1);
              C2 = fmaf(contrib, cc.z, C2);
              A += contrib;
            }
          }
        }
      }
      // every later contribution is scaled by (1-A): below 1e-4 it cannot move an 8-bit channel.  The CTA stops at
      // 5e-5, the same bar that lets this segment hide the ones behind it: stricter than finalize's 1e-4, so rounding
      // in its fold cannot leave a hidden segment visible.
      saturated = __syncthreads_and(!inside || (1.0f - A) < 5e-5f);
      if (saturated) break;
    }
    cp_async_wait<0>();
#ifdef B2S_STATS
    if (threadIdx.x == 0) {
      if (!FIRST) { atomicAdd(&dbg->pad_[0], nchunks); atomicAdd(&dbg->pad_[1], 1); }
      if (FIRST && saturated) atomicAdd(&dbg->pad_[2], 1);
    }
#endif
    if (saturated && threadIdx.x == 0) atomicMin(sat_seg + tile, FIRST ? BS_FIRST - 1 : seg);
    if (FIRST && !saturated && nseg > BS_FIRST && threadIdx.x == 0) {
      // still visible behind the front segments: the rest of the list becomes work items of the second launch
      const int k = nseg - BS_FIRST;
      const int at = atomicAdd(&rest_list[0].x, k);
      for (int q = 0; q < k; ++q) rest_list[1 + at + q] = make_int2(tile, BS_FIRST + q);
    }
    if (nseg > BS_FIRST) {
      float* dst = partial + (size_t)(u0 + seg) * 5 * TILE_PIX + threadIdx.x;
      dst[0] = C0;
      dst[TILE_PIX] = C1;
      dst[2 * TILE_PIX] = C2;
      dst[3 * TILE_PIX] = A;
      if (FIRST) {                                       // segments 1 .. BS_FIRST-1 are part of this result: identity
        for (int q = 1; q < BS_FIRST; ++q) {
          float* z = partial + (size_t)(u0 + q) * 5 * TILE_PIX + threadIdx.x;
          z[0] = 0.f; z[TILE_PIX] = 0.f; z[2 * TILE_PIX] = 0.f; z[3 * TILE_PIX] = 0.f;
        }
      }
      continue;
    }
    // the FIRST kernel saw the whole list: the pixel is final
    if (!inside) continue;
    const size_t p = (size_t)yi * vp.width + xi;
    const float af = fminf(fmaxf(A, 0.0f), 1.0f);
    const float o0 = fminf(fmaxf(C0 + (1.0f - af) * view_bg(vp, 0), 0.0f), 1.0f);
    const float o1 = fminf(fmaxf(C1 + (1.0f - af) * view_bg(vp, 1), 0.0f), 1.0f);
    const float o2 = fminf(fmaxf(C2 + (1.0f - af) * view_bg(vp, 2), 0.0f), 1.0f);
    if (out_rgb != nullptr) {
      out_rgb[3 * p] = o0; out_rgb[3 * p + 1] = o1; out_rgb[3 * p + 2] = o2;
    }
    if (out_alpha != nullptr) out_alpha[p] = af;
    if (out_rgba != nullptr) {   // renderer_cpu.cpp:252-255
      uchar4 q;
      q.x = (unsigned char)(o0 * 255.0f + 0.5f);
      q.y = (unsigned char)(o1 * 255.0f + 0.5f);
      q.z = (unsigned char)(o2 * 255.0f + 0.5f);
      q.w = 255;
      reinterpret_cast<uchar4*>(out_rgba)[p] = q;
    }
  }
  }   // work items
}

// Folds the units of a multi-unit tile front to back: C = C_1 + (1 - A_1) C_2 + ..., A likewise.
__global__ void __launch_bounds__(TILE_PIX)
finalize_sorted_kernel(const ViewParams vp, const int* __restrict__ unit_start, const float* __restrict__ partial,
                       float* __restrict__ out_rgb, float* __restrict__ out_alpha, uint8_t* __restrict__ out_rgba) {
  const int tile = blockIdx.x;
  const int u0 = unit_start[tile], u1 = unit_start[tile + 1];
  if (u1 - u0 <= BS_FIRST) return;                     // the FIRST blend kernel wrote these pixels itself
  const int q = threadIdx.x;
  float C0 = 0.f, C1 = 0.f, C2 = 0.f, A = 0.f;
  for (int u = u0; u < u1; ++u) {
    const float T = 1.0f - A;
    if (T < 1e-4f) break;
    const float* src = partial + (size_t)u * 5 * TILE_PIX + q;
    C0 = fmaf(T, src[0], C0);
    C1 = fmaf(T, src[TILE_PIX], C1);
    C2 = fmaf(T, src[2 * TILE_PIX], C2);
    A = fmaf(T, src[3 * TILE_PIX], A);
  }
  const int xi = (tile % vp.tiles_x) * TILE + (q & 15), yi = (tile / vp.tiles_x) * TILE + (q >> 4);
  if (xi >= vp.width || yi >= vp.height) return;
  const size_t p = (size_t)yi * vp.width + xi;
  const float af = fminf(fmaxf(A, 0.0f), 1.0f);
  const float o0 = fminf(fmaxf(C0 + (1.0f - af) * view_bg(vp, 0), 0.0f), 1.0f);
  const float o1 = fminf(fmaxf(C1 + (1.0f - af) * view_bg(vp, 1), 0.0f), 1.0f);
  const float o2 = fminf(fmaxf(C2 + (1.0f - af) * view_bg(vp, 2), 0.0f), 1.0f);
  if (out_rgb != nullptr) {
    out_rgb[3 * p] = o0; out_rgb[3 * p + 1] = o1; out_rgb[3 * p + 2] = o2;
  }
  if (out_alpha != nullptr) out_alpha[p] = af;
  if (out_rgba != nullptr) {
    uchar4 w;
    w.x = (unsigned char)(o0 * 255.0f + 0.5f);
    w.y = (unsigned char)(o1 * 255.0f + 0.5f);
    w.z = (unsigned char)(o2 * 255.0f + 0.5f);
    w.w = 255;
    reinterpret_cast<uchar4*>(out_rgba)[p] = w;
  }
}

int launch_blend_sorted_fwd(const ViewParams& vp, const float4* rec, const int* vals, const int2* ranges,
                            const int* unit_start, float* partial, int* sat_seg, int2* rest_list, float* out_rgb,
                            float* out_alpha, uint8_t* out_rgba, Counters* dbg, cudaStream_t st) {
  if (vp.n_tiles <= 0) return B2S_OK;
  B2S_CUDA_TRY(cudaMemsetAsync(sat_seg, 0x7f, (size_t)vp.n_tiles * sizeof(int), st));   // "no saturating segment yet"
  B2S_CUDA_TRY(cudaMemsetAsync(rest_list, 0, sizeof(int2), st));                         // no live segments yet
  blend_sorted_units_kernel<true><<<vp.n_tiles, BS_THREADS, 0, st>>>(vp, rec, vals, ranges, unit_start, partial, sat_seg,
                                                                    rest_list, out_rgb, out_alpha, out_rgba, dbg);
  B2S_LAUNCH_CHECK();
  blend_sorted_units_kernel<false><<<sm_count() * 6, BS_THREADS, 0, st>>>(vp, rec, vals, ranges, unit_start, partial, sat_seg,
                                                                         rest_list, out_rgb, out_alpha, out_rgba, dbg);
  B2S_LAUNCH_CHECK();
  finalize_sorted_kernel<<<vp.n_tiles, TILE_PIX, 0, st>>>(vp, unit_start, partial, out_rgb, out_alpha, out_rgba);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

}  // namespace b2s
