// Fit-loop kernels: per-view L1 losses with their image-space gradients, and the fused
// Adam step (+ regulariser gradients).  Reference semantics:
//   python/fit_multiview_stub.py:292-308 (loss), :262,311 (torch.optim.Adam defaults).
// Both are HBM-bound streaming kernels: loss 24..40 B/pixel, Adam 28 B/element.
#include <algorithm>

#include "common.cuh"

namespace b2s {

__device__ __forceinline__ float signf(float v) { return (float)((v > 0.f) - (v < 0.f)); }

__global__ void __launch_bounds__(256)
fit_loss_kernel(const float* __restrict__ rgb, const float* __restrict__ alpha, const float* __restrict__ tgt,
                const float* __restrict__ mask, int hw, float w_sil, float scale, float* __restrict__ g_rgb,
                float* __restrict__ g_alpha, float* __restrict__ loss_accum) {
  const float inv3 = 1.0f / (3.0f * (float)hw), inv1 = 1.0f / (float)hw;
  float local = 0.f;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float d = rgb[3 * (size_t)p + c] - tgt[3 * (size_t)p + c];
      local += fabsf(d) * inv3;
      g_rgb[3 * (size_t)p + c] = scale * inv3 * signf(d);
    }
    if (mask != nullptr) {
      const float d = alpha[p] - mask[p];
      local += w_sil * fabsf(d) * inv1;
      g_alpha[p] = scale * w_sil * inv1 * signf(d);
    }
  }
  __shared__ float ws[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += ws[q];
    atomicAdd(loss_accum, scale * t);
  }
}

int launch_fit_loss(const float* rgb, const float* alpha, const float* tgt, const float* mask, int width,
                    int height, float w_sil, float scale, float* g_rgb, float* g_alpha, float* loss_accum,
                    cudaStream_t st) {
  const int hw = width * height;
  if (hw <= 0) return B2S_OK;
  int blocks = (hw + 255) / 256;
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  fit_loss_kernel<<<blocks, 256, 0, st>>>(rgb, alpha, tgt, mask, hw, w_sil, scale, g_rgb, g_alpha, loss_accum);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

// Target ingestion: 8-bit image planes -> float32 in [0,1] (np.asarray(img, float32) / 255.0,
// python/fit_multiview_stub.py:16-23) on the device, so a fit fed from host memory moves 1 B instead of
// 4 B per value over PCIe.  16 bytes in, 64 bytes out per thread.
__global__ void __launch_bounds__(256)
u8_to_f32_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, long long count) {
  const long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * 16;
  if (i >= count) return;
  if (i + 16 <= count && (reinterpret_cast<uintptr_t>(src + i) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst + i) & 15) == 0) {
    const uint4 v = *reinterpret_cast<const uint4*>(src + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 o;
      o.x = (float)(w[q] & 0xff) / 255.0f;
      o.y = (float)((w[q] >> 8) & 0xff) / 255.0f;
      o.z = (float)((w[q] >> 16) & 0xff) / 255.0f;
      o.w = (float)(w[q] >> 24) / 255.0f;
      reinterpret_cast<float4*>(dst + i)[q] = o;
    }
  } else {
    for (long long k = i; k < count && k < i + 16; ++k) dst[k] = (float)src[k] / 255.0f;
  }
}

int launch_u8_to_f32(const uint8_t* src, float* dst, int64_t count, cudaStream_t st) {
  if (count <= 0) return B2S_OK;
  const long long blocks = (count + 16 * 256 - 1) / (16 * 256);
  u8_to_f32_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, dst, (long long)count);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

// p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)    (torch.optim.Adam, single-tensor form)
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            long long count, float b1, float b2, float eps, float step_size, float inv_sqrt_bc2, long long sb,
            long long se, float reg_s, long long ob, long long oe, float reg_o, const float* __restrict__ skip_flag,
            int* __restrict__ skipped_count) {
  // guard: an iteration whose views overflowed their pair buffers leaves the parameters and moments untouched
  if (skip_flag != nullptr && *skip_flag != 0.0f) {
    if (skipped_count != nullptr && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(skipped_count, 1);
    return;
  }
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    float pi = p[i], gi = g[i];
    if (i >= sb && i < se) gi += reg_s * sigmoidf_acc(pi);                       // d softplus = sigmoid
    if (i >= ob && i < oe) { const float s = sigmoidf_acc(pi); gi += reg_o * s * (1.0f - s); }
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    p[i] = pi - step_size * (mi / denom);
  }
}

// ---- multi-GPU: gradient reduce-scatter + Adam + parameter all-gather in ONE kernel over NVLink multicast (NVLS) -------
// The gradient buffers of the G ranks and their parameter buffers are symmetric allocations bound to a multicast
// address each.  Rank r owns the r-th share of every slice: it PULLS the sum of that share from all ranks with
// multimem.ld_reduce (the NVSwitch adds the G copies and returns one), runs the Adam update of `adam_kernel` on it with
// its own moments, and PUSHES the new parameters to every rank with multimem.st.  The wire carries what an all-reduce
// carries (each gradient crosses once into the switch, each parameter once out of it), Adam costs 1/G, the moments of a
// share live on its owner only, and every replica receives the same bits (one owner per element).  The caller brackets
// the launches of a chunk with two cross-rank barriers (all gradients written before / all parameters landed after).
__device__ __forceinline__ float4 mc_ld_reduce4(const float* mc) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];\n"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc) : "memory");
  return r;
}
__device__ __forceinline__ float mc_ld_reduce1(const float* mc) {
  float r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];\n" : "=f"(r) : "l"(mc) : "memory");
  return r;
}
__device__ __forceinline__ void mc_st4(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};\n" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void mc_st1(float* mc, float v) {
  asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;\n" ::"l"(mc), "f"(v) : "memory");
}

struct AdamConsts {
  float b1, b2, eps, step_size, inv_sqrt_bc2, reg_s, reg_o;
  long long sb, se, ob, oe;
};
__device__ __forceinline__ float adam_one(const AdamConsts& c, long long i, float pi, float gi, float& mi, float& vi) {
  if (i >= c.sb && i < c.se) gi += c.reg_s * sigmoidf_acc(pi);
  if (i >= c.ob && i < c.oe) { const float s = sigmoidf_acc(pi); gi += c.reg_o * s * (1.0f - s); }
  mi = c.b1 * mi + (1.0f - c.b1) * gi;
  vi = c.b2 * vi + (1.0f - c.b2) * gi * gi;
  const float denom = sqrtf(vi) * c.inv_sqrt_bc2 + c.eps;
  return pi - c.step_size * (mi / denom);
}

__global__ void __launch_bounds__(256)
adam_multimem_kernel(float* __restrict__ p_mc, const float* __restrict__ g_mc, const float* __restrict__ p_loc,
                     float* __restrict__ m, float* __restrict__ v, long long lo, long long hi, AdamConsts c,
                     const float* __restrict__ skip_flag, int* __restrict__ skipped_count) {
  if (skip_flag != nullptr && *skip_flag != 0.0f) {      // the same reduced value on every rank: all skip or none
    if (skipped_count != nullptr && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(skipped_count, 1);
    return;
  }
  const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n4 = (hi - lo) / 4;
  for (long long q = t0; q < n4; q += stride) {
    const long long i = lo + 4 * q;
    const float4 g = mc_ld_reduce4(g_mc + i);
    const float4 pv = *reinterpret_cast<const float4*>(p_loc + i);
    float4 mv = *reinterpret_cast<const float4*>(m + i), vv = *reinterpret_cast<const float4*>(v + i);
    float4 pn;
    pn.x = adam_one(c, i, pv.x, g.x, mv.x, vv.x);
    pn.y = adam_one(c, i + 1, pv.y, g.y, mv.y, vv.y);
    pn.z = adam_one(c, i + 2, pv.z, g.z, mv.z, vv.z);
    pn.w = adam_one(c, i + 3, pv.w, g.w, mv.w, vv.w);
    *reinterpret_cast<float4*>(m + i) = mv;
    *reinterpret_cast<float4*>(v + i) = vv;
    mc_st4(p_mc + i, pn);
  }
  for (long long i = lo + 4 * n4 + t0; i < hi; i += stride) {     // at most 3 elements at the end of a slice
    const float g = mc_ld_reduce1(g_mc + i);
    float mi = m[i], vi = v[i];
    const float pn = adam_one(c, i, p_loc[i], g, mi, vi);
    m[i] = mi;
    v[i] = vi;
    mc_st1(p_mc + i, pn);
  }
}

__global__ void __launch_bounds__(64) tail_multimem_kernel(const float* __restrict__ tail_mc, float* __restrict__ out, int count) {
  if ((int)threadIdx.x < count) out[threadIdx.x] = mc_ld_reduce1(tail_mc + threadIdx.x);
}

// rank r's share [lo, hi) of a slice of `count` floats: ceil(count / world) rounded up to 4 floats (16-byte multimem
// accesses), clipped to the slice -- the shares tile the slice, every boundary but the last is a multiple of 4
void multimem_share(int64_t count, int rank, int world, long long* lo, long long* hi) {
  const long long per = (((long long)count + world - 1) / world + 3) / 4 * 4;
  *lo = std::min<long long>((long long)rank * per, count);
  *hi = std::min<long long>(*lo + per, count);
}

int launch_adam_multimem(float* params_mc, const float* grads_mc, const float* params_local, float* m, float* v,
                         int64_t count, int rank, int world, int step, float lr, float b1, float b2, float eps, int64_t sb,
                         int64_t se, float reg_scale, int64_t ob, int64_t oe, float reg_op, const float* skip_flag,
                         int* skipped_count, cudaStream_t st) {
  if (count <= 0 || world <= 0) return B2S_OK;
  const double bc1 = 1.0 - pow((double)b1, (double)step);
  const double bc2 = 1.0 - pow((double)b2, (double)step);
  AdamConsts c;
  c.b1 = b1; c.b2 = b2; c.eps = eps;
  c.step_size = (float)((double)lr / bc1);
  c.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  c.reg_s = (se > sb) ? reg_scale / (float)(se - sb) : 0.f;
  c.reg_o = (oe > ob) ? reg_op / (float)(oe - ob) : 0.f;
  c.sb = sb; c.se = se; c.ob = ob; c.oe = oe;
  long long lo, hi;
  multimem_share(count, rank, world, &lo, &hi);
  if (hi <= lo) return B2S_OK;
  long long blocks = ((hi - lo) / 4 + 255) / 256 + 1;
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  adam_multimem_kernel<<<(int)blocks, 256, 0, st>>>(params_mc, grads_mc, params_local, m, v, lo, hi, c, skip_flag, skipped_count);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int launch_tail_multimem(const float* tail_mc, float* out, int count, cudaStream_t st) {
  if (count <= 0 || count > 64) { set_error("tail of %d floats (1..64)", count); return B2S_ERR_INVALID; }
  tail_multimem_kernel<<<1, 64, 0, st>>>(tail_mc, out, count);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int launch_adam(float* params, const float* grads, float* m, float* v, int64_t count, int step, float lr,
                float b1, float b2, float eps, int64_t sb, int64_t se, float reg_scale, int64_t ob, int64_t oe,
                float reg_op, const float* skip_flag, int* skipped_count, cudaStream_t st) {
  if (count <= 0) return B2S_OK;
  const double bc1 = 1.0 - pow((double)b1, (double)step);
  const double bc2 = 1.0 - pow((double)b2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  const float rs = (se > sb) ? reg_scale / (float)(se - sb) : 0.f;
  const float ro = (oe > ob) ? reg_op / (float)(oe - ob) : 0.f;
  long long blocks = (count + 255) / 256;
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  adam_kernel<<<(int)blocks, 256, 0, st>>>(params, grads, m, v, (long long)count, b1, b2, eps, step_size,
                                            inv_sqrt_bc2, (long long)sb, (long long)se, rs, (long long)ob,
                                            (long long)oe, ro, skip_flag, skipped_count);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

}  // namespace b2s
