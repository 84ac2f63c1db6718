// Forward blending.  The reference scatters 4 global atomicAdds per (Gaussian,pixel) pair
// (src/renderer.cu:98-102) or runs 10+ elementwise passes over (256,H,W) temporaries
// (python/torch_renderer.py:167-190); here every pixel accumulates in registers.
//
//   WSUM  : A += w c ; W += w ; D += w z ; out = clamp((bg+A)/(1+W))   torch_renderer.py:181-202
//   SORTED: front-to-back "over" with per-pixel alpha state           renderer_cpu.cpp:196-215,241-257
//
// WSUM schedule (v2): the work unit is (tile, segment of <= SEG Gaussians); ONE WARP owns a unit
// and every lane owns 8 pixels (a column of 8 rows), so a Gaussian record is read from shared
// memory once per 256 pixel-pairs and the x-term of the exponent is shared by the 8 rows.  The
// unit's records are gathered with cp.async into a per-warp 3-stage ring (no block barriers).
// Because the weighted sum is order independent, a tile whose list spans several units is
// summed from per-unit partial accumulators by finalize_kernel (fixed order: deterministic).
// Bound: FP32 issue + MUFU.EX2 -- 8 FP32 + 1 ex2 per pixel-pair (+1 FFMA with depth).
#include "common.cuh"

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
  const float o0 = fminf(fmaxf((vp.bg[0] + R) * inv, 0.0f), 1.0f);
  const float o1 = fminf(fmaxf((vp.bg[1] + G) * inv, 0.0f), 1.0f);
  const float o2 = fminf(fmaxf((vp.bg[2] + B) * inv, 0.0f), 1.0f);
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
  float4 a[FW_CHUNK];
  float4 b[FW_CHUNK];
  float4 c[FW_CHUNK];
};

// EXACT=false : w = 2^(qx dx^2 + qy dy^2 + log2 op) on every pixel of the tile
// EXACT=true  : w = op * 2^(...), only inside the Gaussian's pixel bbox and if w >= 1e-5
//               (renderer_cpu.cpp:107-113 -- the native weighted-sum mode)
template <bool DEPTH, bool EXACT>
__global__ void __launch_bounds__(FW_WARPS * 32)
blend_wsum_fwd_kernel(const ViewParams vp, const float4* __restrict__ rec, const int* __restrict__ vals,
                      const int2* __restrict__ ranges, const int* __restrict__ unit_start,
                      const int2* __restrict__ units, float* __restrict__ partial, float* __restrict__ out_rgb,
                      float* __restrict__ out_alpha, float* __restrict__ out_depth, float* __restrict__ acc,
                      uint8_t* __restrict__ out_rgba) {
  __shared__ __align__(16) FwStage ring[FW_WARPS][FW_STAGES];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u = blockIdx.x * FW_WARPS + warp;
  if (u >= unit_start[vp.n_tiles]) return;           // warps are independent: no block barrier below
  const int2 ud = units[u];
  const int tile = ud.x;
  const int2 rg = ranges[tile];
  const int start = rg.x + ud.y * SEG;
  const int n = max(0, min(SEG, rg.y - start));
  const int nseg = unit_start[tile + 1] - unit_start[tile];
  const int tx = tile % vp.tiles_x, ty = tile / vp.tiles_x;
  const int cx = lane & 15, half = lane >> 4;
  const int xi = tx * TILE + cx, yi0 = ty * TILE + half * FW_ROWS;
  const float x = xi + 0.5f, y0 = yi0 + 0.5f;
  FwStage* my = ring[warp];

  float R[FW_ROWS], G[FW_ROWS], B[FW_ROWS], W[FW_ROWS], D[FW_ROWS];
#pragma unroll
  for (int r = 0; r < FW_ROWS; ++r) R[r] = G[r] = B[r] = W[r] = D[r] = 0.f;

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
      const float4 a = s.a[j];
      const float4 b = s.b[j];
      const float dx = x - a.x;
      const float dy0 = y0 - a.y;
      float zz = 0.f;
      if (DEPTH) zz = s.c[j].x;
      if constexpr (!EXACT) {
        const float tx2 = fmaf(a.z * dx, dx, b.w);
#pragma unroll
        for (int r = 0; r < FW_ROWS; ++r) {
          const float dy = dy0 + (float)r;
          const float w = ex2_approx(fmaf(a.w * dy, dy, tx2));
          W[r] += w;
          R[r] = fmaf(w, b.x, R[r]);
          G[r] = fmaf(w, b.y, G[r]);
          B[r] = fmaf(w, b.z, B[r]);
          if (DEPTH) D[r] = fmaf(w, zz, D[r]);
        }
      } else {
        const float4 cc = s.c[j];
        const float tx2 = a.z * dx * dx;
        const int bx = __float_as_int(cc.y), by = __float_as_int(cc.z);
        const bool inx = (xi >= (bx & 0xffff)) && (xi <= (bx >> 16));
        const int ymin = by & 0xffff, ymax = by >> 16;
#pragma unroll
        for (int r = 0; r < FW_ROWS; ++r) {
          const float dy = dy0 + (float)r;
          float w = b.w * ex2_approx(fmaf(a.w * dy, dy, tx2));
          w = (inx && (yi0 + r) >= ymin && (yi0 + r) <= ymax && w >= 1e-5f) ? w : 0.0f;
          W[r] += w;
          R[r] = fmaf(w, b.x, R[r]);
          G[r] = fmaf(w, b.y, G[r]);
          B[r] = fmaf(w, b.z, B[r]);
          if (DEPTH) D[r] = fmaf(w, zz, D[r]);
        }
      }
    }
    __syncwarp();
  }
  cp_async_wait<0>();

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
                          const int* unit_start, const int2* units, int64_t unit_cap, float* partial,
                          float* out_rgb, float* out_alpha, float* out_depth, float* acc, uint8_t* out_rgba,
                          cudaStream_t st) {
  if (vp.n_tiles <= 0) return B2S_OK;
  const int blocks = (int)((unit_cap + FW_WARPS - 1) / FW_WARPS);
  const bool depth = out_depth != nullptr;   // D is only accumulated when the depth image is requested
#define B2S_FW(DD, EE)                                                                                              \
  blend_wsum_fwd_kernel<DD, EE><<<blocks, FW_WARPS * 32, 0, st>>>(vp, rec, vals, ranges, unit_start, units, partial, \
                                                                   out_rgb, out_alpha, out_depth, acc, out_rgba)
  if (vp.exact_bbox) { if (depth) B2S_FW(true, true); else B2S_FW(false, true); }
  else               { if (depth) B2S_FW(true, false); else B2S_FW(false, false); }
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
// One CTA per tile, one pixel per thread, per-pixel alpha state in a register, block-wide early
// termination once every pixel of the tile is saturated.  Order matters, so a tile's list is
// NOT split into units here.
constexpr int BS_THREADS = 256;
constexpr int BS_CHUNK = 128;

struct StageBuf {
  float4 a[BS_CHUNK];
  float4 b[BS_CHUNK];
  float4 c[BS_CHUNK];
};

__device__ __forceinline__ void stage_chunk(StageBuf& sb, const float4* __restrict__ rec, const int* __restrict__ vals,
                                            int start, int n, int chunk) {
  for (int t = threadIdx.x; t < BS_CHUNK; t += BS_THREADS) {
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

__global__ void __launch_bounds__(BS_THREADS)
blend_sorted_fwd_kernel(const ViewParams vp, const float4* __restrict__ rec, const int* __restrict__ vals,
                        const int2* __restrict__ ranges, float* __restrict__ out_rgb, float* __restrict__ out_alpha,
                        uint8_t* __restrict__ out_rgba) {
  __shared__ __align__(16) StageBuf sb[2];
  const int tile = blockIdx.x;
  const int tx = tile % vp.tiles_x, ty = tile / vp.tiles_x;
  const int xi = tx * TILE + (threadIdx.x & 15), yi = ty * TILE + (threadIdx.x >> 4);
  const float x = xi + 0.5f, y = yi + 0.5f;
  const int2 rg = ranges[tile];
  const int n = rg.y - rg.x;
  const int nchunks = (n + BS_CHUNK - 1) / BS_CHUNK;
  float C0 = 0.f, C1 = 0.f, C2 = 0.f, A = 0.f;

  if (nchunks > 0) stage_chunk(sb[0], rec, vals, rg.x, n, 0);
  cp_async_commit();
  for (int c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) stage_chunk(sb[(c + 1) & 1], rec, vals, rg.x, n, c + 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const StageBuf& s = sb[c & 1];
    const int cnt = min(BS_CHUNK, n - c * BS_CHUNK);
#pragma unroll 4
    for (int j = 0; j < cnt; ++j) {
      const float4 a = s.a[j];
      const float4 b = s.b[j];
      const float4 cc = s.c[j];
      const float dx = x - a.x, dy = y - a.y;
      float al = b.w * ex2_approx(fmaf(a.w * dy, dy, a.z * dx * dx));
      const int bx = __float_as_int(cc.y), by = __float_as_int(cc.z);
      const bool in = (xi >= (bx & 0xffff)) && (xi <= (bx >> 16)) && (yi >= (by & 0xffff)) && (yi <= (by >> 16));
      if (in && al >= 1e-5f) {
        al = fminf(al, 1.0f);
        const float contrib = (1.0f - A) * al;
        if (contrib > 0.0f) {
          C0 = fmaf(contrib, b.x, C0);
          C1 = fmaf(contrib, b.y, C1);
          C2 = fmaf(contrib, b.z, C2);
          A += contrib;
        }
      }
    }
    // every later contribution is scaled by (1-A): below 1e-4 it cannot move an 8-bit channel
    const int done = (1.0f - A) < 1e-4f;
    if (__syncthreads_and(done)) break;
  }
  cp_async_wait<0>();
  if (xi >= vp.width || yi >= vp.height) return;
  const size_t p = (size_t)yi * vp.width + xi;
  const float af = fminf(fmaxf(A, 0.0f), 1.0f);
  const float o0 = fminf(fmaxf(C0 + (1.0f - af) * vp.bg[0], 0.0f), 1.0f);
  const float o1 = fminf(fmaxf(C1 + (1.0f - af) * vp.bg[1], 0.0f), 1.0f);
  const float o2 = fminf(fmaxf(C2 + (1.0f - af) * vp.bg[2], 0.0f), 1.0f);
  if (out_rgb != nullptr) {
    out_rgb[3 * p] = o0; out_rgb[3 * p + 1] = o1; out_rgb[3 * p + 2] = o2;
  }
  if (out_alpha != nullptr) out_alpha[p] = af;
  if (out_rgba != nullptr) {   // renderer_cpu.cpp:252-255
    uchar4 u;
    u.x = (unsigned char)(o0 * 255.0f + 0.5f);
    u.y = (unsigned char)(o1 * 255.0f + 0.5f);
    u.z = (unsigned char)(o2 * 255.0f + 0.5f);
    u.w = 255;
    reinterpret_cast<uchar4*>(out_rgba)[p] = u;
  }
}

int launch_blend_sorted_fwd(const ViewParams& vp, const float4* rec, const int* vals, const int2* ranges,
                            float* out_rgb, float* out_alpha, uint8_t* out_rgba, cudaStream_t st) {
  if (vp.n_tiles <= 0) return B2S_OK;
  blend_sorted_fwd_kernel<<<vp.n_tiles, BS_THREADS, 0, st>>>(vp, rec, vals, ranges, out_rgb, out_alpha, out_rgba);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

}  // namespace b2s
