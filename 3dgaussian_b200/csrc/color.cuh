// Activations and the colour model (SH basis, its gradient, coefficient loads) shared by the per-Gaussian kernels of
// preprocess.cu (reference projection) and splat2d.cu (extension modes).
// Reference: python/torch_renderer.py:86-106 (`_eval_colors`), python/fit_multiview_stub.py:269-274 (activations).
#pragma once
#include "common.cuh"

namespace b2s {

__device__ __forceinline__ float act_scale(const ViewParams& vp, float raw) {
  return (vp.act & B2S_ACT_SCALES_SOFTPLUS) ? softplusf_acc(raw) + 1e-3f : raw;
}
__device__ __forceinline__ float act_opac(const ViewParams& vp, float raw) {
  return (vp.act & B2S_ACT_OPACITY_SIGMOID) ? sigmoidf_acc(raw) : raw;
}

// Basis of the colour model: k=0..3 is the reference's [1, d.x, d.y, d.z]
// (torch_renderer.py:98-103); k=4..15 are standard real-SH band 2/3 polynomials (extension).
__device__ __forceinline__ void sh_basis(float x, float y, float z, int K, float* b) {
  b[0] = 1.0f; b[1] = x; b[2] = y; b[3] = z;
  if (K > 4) {
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    b[4] = 1.0925484305920792f * xy;
    b[5] = -1.0925484305920792f * yz;
    b[6] = 0.31539156525252005f * (2.0f * zz - xx - yy);
    b[7] = -1.0925484305920792f * xz;
    b[8] = 0.5462742152960396f * (xx - yy);
    if (K > 9) {
      b[9] = -0.5900435899266435f * y * (3.0f * xx - yy);
      b[10] = 2.890611442640554f * xy * z;
      b[11] = -0.4570457994644658f * y * (4.0f * zz - xx - yy);
      b[12] = 0.3731763325901154f * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
      b[13] = -0.4570457994644658f * x * (4.0f * zz - xx - yy);
      b[14] = 1.445305721320277f * z * (xx - yy);
      b[15] = -0.5900435899266435f * x * (xx - 3.0f * yy);
    }
  }
}

// d basis_k / d (x,y,z)
__device__ __forceinline__ void sh_basis_grad(float x, float y, float z, int K, float* bx, float* by, float* bz) {
  bx[0] = by[0] = bz[0] = 0.0f;
  bx[1] = 1.0f; by[1] = 0.0f; bz[1] = 0.0f;
  bx[2] = 0.0f; by[2] = 1.0f; bz[2] = 0.0f;
  bx[3] = 0.0f; by[3] = 0.0f; bz[3] = 1.0f;
  if (K > 4) {
    const float c0 = 1.0925484305920792f, c2 = 0.31539156525252005f, c4 = 0.5462742152960396f;
    bx[4] = c0 * y;          by[4] = c0 * x;          bz[4] = 0.0f;
    bx[5] = 0.0f;            by[5] = -c0 * z;         bz[5] = -c0 * y;
    bx[6] = -2.0f * c2 * x;  by[6] = -2.0f * c2 * y;  bz[6] = 4.0f * c2 * z;
    bx[7] = -c0 * z;         by[7] = 0.0f;            bz[7] = -c0 * x;
    bx[8] = 2.0f * c4 * x;   by[8] = -2.0f * c4 * y;  bz[8] = 0.0f;
    if (K > 9) {
      const float xx = x * x, yy = y * y, zz = z * z;
      const float d0 = -0.5900435899266435f, d1 = 2.890611442640554f, d2 = -0.4570457994644658f,
                  d3 = 0.3731763325901154f, d5 = 1.445305721320277f;
      // b9 = d0*y*(3xx-yy)
      bx[9] = d0 * 6.0f * x * y;  by[9] = d0 * (3.0f * xx - 3.0f * yy);  bz[9] = 0.0f;
      // b10 = d1*x*y*z
      bx[10] = d1 * y * z;  by[10] = d1 * x * z;  bz[10] = d1 * x * y;
      // b11 = d2*y*(4zz-xx-yy)
      bx[11] = d2 * (-2.0f * x * y);  by[11] = d2 * (4.0f * zz - xx - 3.0f * yy);  bz[11] = d2 * 8.0f * y * z;
      // b12 = d3*z*(2zz-3xx-3yy)
      bx[12] = d3 * (-6.0f * x * z);  by[12] = d3 * (-6.0f * y * z);  bz[12] = d3 * (6.0f * zz - 3.0f * xx - 3.0f * yy);
      // b13 = d2*x*(4zz-xx-yy)
      bx[13] = d2 * (4.0f * zz - 3.0f * xx - yy);  by[13] = d2 * (-2.0f * x * y);  bz[13] = d2 * 8.0f * x * z;
      // b14 = d5*z*(xx-yy)
      bx[14] = d5 * 2.0f * x * z;  by[14] = d5 * (-2.0f * y * z);  bz[14] = d5 * (xx - yy);
      // b15 = d0*x*(xx-3yy)
      bx[15] = d0 * (3.0f * xx - 3.0f * yy);  by[15] = d0 * (-6.0f * x * y);  bz[15] = 0.0f;
    }
  }
}

template <int K>
__device__ __forceinline__ void load_coeffs(const float* __restrict__ colors, int i, float* c) {
  // (N,K,3) row-major: 3K contiguous floats per Gaussian; 12K bytes is a multiple of 16 when K%4==0
  if constexpr ((K * 3) % 4 == 0) {
    const float4* p = reinterpret_cast<const float4*>(colors + (size_t)i * K * 3);
#pragma unroll
    for (int q = 0; q < K * 3 / 4; ++q) {
      const float4 v = __ldg(p + q);
      c[4 * q] = v.x; c[4 * q + 1] = v.y; c[4 * q + 2] = v.z; c[4 * q + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int q = 0; q < K * 3; ++q) c[q] = __ldg(colors + (size_t)i * K * 3 + q);
  }
}

// raw (pre-clamp) colour + the unit direction used by the SH model
template <int K>
__device__ __forceinline__ void eval_color(const ViewParams& vp, const float* c, float mx, float my, float mz,
                                           float* rgb_raw, float* dir, float* rinv) {
  if constexpr (K == 1) {
    rgb_raw[0] = c[0]; rgb_raw[1] = c[1]; rgb_raw[2] = c[2];
    if (vp.act & B2S_ACT_COLORS_SIGMOID) {
#pragma unroll
      for (int q = 0; q < 3; ++q) rgb_raw[q] = sigmoidf_acc(rgb_raw[q]);
    }
    dir[0] = dir[1] = dir[2] = 0.0f;
    *rinv = 0.0f;
  } else {
    const float vx = vp.cam[0] - mx, vy = vp.cam[1] - my, vz = vp.cam[2] - mz;
    const float r = sqrtf(vx * vx + vy * vy + vz * vz);
    const float inv = 1.0f / (r + 1e-8f);
    dir[0] = vx * inv; dir[1] = vy * inv; dir[2] = vz * inv;
    *rinv = inv;
    float b[16];
    sh_basis(dir[0], dir[1], dir[2], K, b);
    float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      r0 = fmaf(b[k], c[3 * k], r0);
      r1 = fmaf(b[k], c[3 * k + 1], r1);
      r2 = fmaf(b[k], c[3 * k + 2], r2);
    }
    rgb_raw[0] = r0; rgb_raw[1] = r1; rgb_raw[2] = r2;
  }
}

constexpr float NEG_HALF_LOG2E = -0.72134752044448170368f;  // -0.5 * log2(e)

}  // namespace b2s
