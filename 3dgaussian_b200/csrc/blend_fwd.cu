// Forward blending: one CTA per 16x16 tile, the tile's Gaussian records are staged through
// shared memory with cp.async (double buffered) and every thread accumulates its pixels in
// registers -- no atomics (the reference scatters 4 global atomicAdds per pair,
// src/renderer.cu:98-102, or runs 10+ elementwise passes over (256,H,W) temporaries,
// python/torch_renderer.py:167-190).
//
//   WSUM  : A += w c ; W += w ; D += w z ; out = clamp((bg+A)/(1+W))   torch_renderer.py:181-202
//   SORTED: front-to-back "over" with per-pixel alpha state           renderer_cpu.cpp:196-215,241-257
//
// Bound: FP32 issue + MUFU.EX2 (8 FP32 + 1 ex2 per pixel-pair); HBM traffic is 48 B per
// (Gaussian,tile) pair + 20..40 B per pixel.
#include "common.cuh"

namespace b2s {

constexpr int BF_THREADS = 128;   // 2 pixels per thread: (cx, r) and (cx, r+8)
constexpr int BF_CHUNK = 128;     // Gaussians staged per buffer

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

struct StageBuf {
  float4 a[BF_CHUNK];
  float4 b[BF_CHUNK];
  float4 c[BF_CHUNK];
};

template <int THREADS>
__device__ __forceinline__ void stage_chunk(StageBuf& sb, const float4* __restrict__ rec, const int* __restrict__ vals,
                                            int start, int n, int chunk) {
  for (int t = threadIdx.x; t < BF_CHUNK; t += THREADS) {
    const int i = chunk * BF_CHUNK + t;
    if (i < n) {
      const int id = __ldg(vals + start + i);
      const float4* src = rec + 3 * (size_t)id;
      cp_async16(&sb.a[t], src);
      cp_async16(&sb.b[t], src + 1);
      cp_async16(&sb.c[t], src + 2);
    }
  }
}

// EXACT=false : w = 2^(qx dx^2 + qy dy^2 + log2 op), evaluated on every pixel of the tile
// EXACT=true  : w = op * 2^(...), restricted to the Gaussian's pixel bbox and w >= 1e-5
//               (renderer_cpu.cpp:107-113 -- the native weighted-sum mode)
template <bool EXACT>
__global__ void __launch_bounds__(BF_THREADS)
blend_wsum_fwd_kernel(const ViewParams vp, const float4* __restrict__ rec, const int* __restrict__ vals,
                      const int2* __restrict__ ranges, float* __restrict__ out_rgb, float* __restrict__ out_alpha,
                      float* __restrict__ out_depth, float* __restrict__ acc, uint8_t* __restrict__ out_rgba) {
  __shared__ __align__(16) StageBuf sb[2];
  const int tile = blockIdx.x;
  const int tx = tile % vp.tiles_x, ty = tile / vp.tiles_x;
  const int cx = threadIdx.x & 15, r = threadIdx.x >> 4;
  const int xi = tx * TILE + cx, yi0 = ty * TILE + r, yi1 = yi0 + 8;
  const float x = xi + 0.5f, y0 = yi0 + 0.5f, y1 = yi1 + 0.5f;
  const int2 rg = ranges[tile];
  const int n = rg.y - rg.x;
  const int nchunks = (n + BF_CHUNK - 1) / BF_CHUNK;

  float R0 = 0.f, G0 = 0.f, B0 = 0.f, W0 = 0.f, D0 = 0.f;
  float R1 = 0.f, G1 = 0.f, B1 = 0.f, W1 = 0.f, D1 = 0.f;

  if (nchunks > 0) stage_chunk<BF_THREADS>(sb[0], rec, vals, rg.x, n, 0);
  cp_async_commit();
  for (int c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) stage_chunk<BF_THREADS>(sb[(c + 1) & 1], rec, vals, rg.x, n, c + 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const StageBuf& s = sb[c & 1];
    const int cnt = min(BF_CHUNK, n - c * BF_CHUNK);
#pragma unroll 4
    for (int j = 0; j < cnt; ++j) {
      const float4 a = s.a[j];
      const float4 b = s.b[j];
      const float4 cc = s.c[j];
      const float dx = x - a.x;
      const float dy0 = y0 - a.y, dy1 = y1 - a.y;
      float w0, w1;
      if constexpr (!EXACT) {
        const float tx2 = fmaf(a.z * dx, dx, b.w);
        w0 = ex2_approx(fmaf(a.w * dy0, dy0, tx2));
        w1 = ex2_approx(fmaf(a.w * dy1, dy1, tx2));
      } else {
        const float tx2 = a.z * dx * dx;
        w0 = b.w * ex2_approx(fmaf(a.w * dy0, dy0, tx2));
        w1 = b.w * ex2_approx(fmaf(a.w * dy1, dy1, tx2));
        const int bx = __float_as_int(cc.y), by = __float_as_int(cc.z);
        const bool inx = (xi >= (bx & 0xffff)) && (xi <= (bx >> 16));
        const int ymin = by & 0xffff, ymax = by >> 16;
        w0 = (inx && yi0 >= ymin && yi0 <= ymax && w0 >= 1e-5f) ? w0 : 0.0f;
        w1 = (inx && yi1 >= ymin && yi1 <= ymax && w1 >= 1e-5f) ? w1 : 0.0f;
      }
      W0 += w0; W1 += w1;
      R0 = fmaf(w0, b.x, R0); R1 = fmaf(w1, b.x, R1);
      G0 = fmaf(w0, b.y, G0); G1 = fmaf(w1, b.y, G1);
      B0 = fmaf(w0, b.z, B0); B1 = fmaf(w1, b.z, B1);
      D0 = fmaf(w0, cc.x, D0); D1 = fmaf(w1, cc.x, D1);
    }
    __syncthreads();
  }
  cp_async_wait<0>();

  const size_t hw = (size_t)vp.width * vp.height;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int yi = q ? yi1 : yi0;
    if (xi >= vp.width || yi >= vp.height) continue;
    const float R = q ? R1 : R0, G = q ? G1 : G0, B = q ? B1 : B0, W = q ? W1 : W0, D = q ? D1 : D0;
    const size_t p = (size_t)yi * vp.width + xi;
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
    if (out_rgba != nullptr) {   // renderer_cpu.cpp:236-239 quantisation
      uchar4 u;
      u.x = (unsigned char)(o0 * 255.0f + 0.5f);
      u.y = (unsigned char)(o1 * 255.0f + 0.5f);
      u.z = (unsigned char)(o2 * 255.0f + 0.5f);
      u.w = 255;
      reinterpret_cast<uchar4*>(out_rgba)[p] = u;
    }
  }
}

int launch_blend_wsum_fwd(const ViewParams& vp, const float4* rec, const int* vals, const int2* ranges,
                          float* out_rgb, float* out_alpha, float* out_depth, float* acc, uint8_t* out_rgba,
                          cudaStream_t st) {
  if (vp.n_tiles <= 0) return B2S_OK;
  if (vp.exact_bbox)
    blend_wsum_fwd_kernel<true><<<vp.n_tiles, BF_THREADS, 0, st>>>(vp, rec, vals, ranges, out_rgb, out_alpha, out_depth, acc, out_rgba);
  else
    blend_wsum_fwd_kernel<false><<<vp.n_tiles, BF_THREADS, 0, st>>>(vp, rec, vals, ranges, out_rgb, out_alpha, out_depth, acc, out_rgba);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

// Depth-sorted "over" compositing, one pixel per thread, per-pixel alpha state in a register,
// block-wide early termination once every pixel of the tile is saturated.
constexpr int BS_THREADS = 256;

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
  const int nchunks = (n + BF_CHUNK - 1) / BF_CHUNK;
  float C0 = 0.f, C1 = 0.f, C2 = 0.f, A = 0.f;

  if (nchunks > 0) stage_chunk<BS_THREADS>(sb[0], rec, vals, rg.x, n, 0);
  cp_async_commit();
  for (int c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) stage_chunk<BS_THREADS>(sb[(c + 1) & 1], rec, vals, rg.x, n, c + 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const StageBuf& s = sb[c & 1];
    const int cnt = min(BF_CHUNK, n - c * BF_CHUNK);
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
