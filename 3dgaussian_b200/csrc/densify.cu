// Densify / prune compaction on the device.
// Reference semantics: python/fit_multiview_stub.py:140-197 (_densify_and_prune)
//   keep   = sigmoid(op_raw) > prune_opacity            (fewer than 64 kept -> keep the top 64 by opacity)   :153-157
//   filter = order-preserving                                                                                  :159-163
//   add_n  = min(max_gaussians - n, int(n * densify_ratio))                                                   :165-167
//   clones = the add_n most opaque kept Gaussians: mean + 0.25 * scale * N(0,1), same scales_raw,
//            op_raw - 0.1, same colours / SH                                                                    :169-195
// The reference does this with torch indexing + topk + randn_like and then builds a NEW Adam optimizer
// (state reset, :319-325).  Here: one sortable key per Gaussian, an 8-bit x 4-pass radix SELECT for the
// k-th largest opacity (ties resolved by lowest index), two exclusive scans and one scatter kernel that
// writes survivors (stable) followed by the clones (in source order; torch.topk orders them by value --
// the SET is the same).  Jitter is counter-based Philox4x32-10 keyed by (seed, iteration) with the SOURCE
// index as counter, so every rank of a multi-GPU fit produces identical clones without a broadcast.
// Nothing here is on the per-iteration path (it runs every ~100 iterations); it exists so that the fit
// never leaves the device.
#include "common.cuh"

namespace b2s {

struct DpState {       // device-side control block
  int count0;          // sigmoid(op) > threshold
  int use_top64;       // fewer than 64 kept -> top-64 mode
  int n1;              // survivors
  int add_n;           // clones
  // radix-select state
  unsigned prefix;     // threshold key so far
  int k_rem;           // how many still to take among keys matching the prefix
  int active;          // 0: select nothing
  int pad;
  int hist[256];
};

__device__ __forceinline__ unsigned sortable_key(float v) {   // ascending unsigned order == ascending float order
  const unsigned u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(256)
dp_keys_kernel(int n, const float* __restrict__ op_raw, float thr, unsigned* __restrict__ keys, int* __restrict__ keep0,
               DpState* __restrict__ S) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  int k = 0;
  if (i < n) {
    const float op = sigmoidf_acc(op_raw[i]);
    keys[i] = sortable_key(op);
    k = op > thr ? 1 : 0;
    keep0[i] = k;
  }
  const int c = __syncthreads_count(k);
  if (threadIdx.x == 0 && c) atomicAdd(&S->count0, c);
}

// phase 0: decide the prune mode and arm the top-64 select; phase 1: survivors -> add_n, arm the clone select
__global__ void dp_decide_kernel(DpState* S, int phase, int n, int max_gaussians, double ratio, const int* n1_from_scan) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (phase == 0) {
    S->use_top64 = S->count0 < 64 ? 1 : 0;
    S->prefix = 0u;
    S->k_rem = n < 64 ? n : 64;
    S->active = S->use_top64;
  } else {
    const int n1 = *n1_from_scan;
    S->n1 = n1;
    const int room = max_gaussians - n1 > 0 ? max_gaussians - n1 : 0;
    int want = (int)((double)n1 * ratio);         // int(n * densify_ratio), Python float = double
    if (want < 0) want = 0;
    int add = room < want ? room : want;
    if (add > n1) add = n1;                        // topk(k = min(n, add_n))
    S->add_n = add;
    S->prefix = 0u;
    S->k_rem = add;
    S->active = add > 0 ? 1 : 0;
  }
  for (int d = 0; d < 256; ++d) S->hist[d] = 0;
}

// histogram of the digit at `shift` over candidates whose higher digits equal the prefix
__global__ void __launch_bounds__(256)
dp_hist_kernel(int n, const unsigned* __restrict__ keys, const int* __restrict__ cand, DpState* __restrict__ S, int shift) {
  __shared__ int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  if (S->active) {
    const unsigned prefix = S->prefix;
    const unsigned himask = shift >= 24 ? 0u : (0xffffffffu << (shift + 8));
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
      if (cand != nullptr && !cand[i]) continue;
      const unsigned k = keys[i];
      if ((k & himask) == (prefix & himask)) atomicAdd(&h[(k >> shift) & 255u], 1);
    }
  }
  __syncthreads();
  if (h[threadIdx.x]) atomicAdd(&S->hist[threadIdx.x], h[threadIdx.x]);
}

// walk the digits from the top until k_rem is covered
__global__ void dp_pick_kernel(DpState* S, int shift) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (S->active) {
    int k = S->k_rem, d = 255;
    for (; d > 0; --d) {
      if (S->hist[d] >= k) break;
      k -= S->hist[d];
    }
    S->prefix |= (unsigned)d << shift;
    S->k_rem = k;      // still to take among keys with this digit (at the last pass: ties at the threshold)
  }
  for (int q = 0; q < 256; ++q) S->hist[q] = 0;
}

// gt[i] = candidate above the threshold key, eq[i] = candidate equal to it
__global__ void __launch_bounds__(256)
dp_mark_kernel(int n, const unsigned* __restrict__ keys, const int* __restrict__ cand, const DpState* __restrict__ S,
               int* __restrict__ gt, int* __restrict__ eq) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const bool c = S->active && (cand == nullptr || cand[i]);
  const unsigned k = keys[i], T = S->prefix;
  gt[i] = (c && k > T) ? 1 : 0;
  eq[i] = (c && k == T) ? 1 : 0;
}

// sel[i] = gt || (eq && rank among the equals < k_rem); eq_rank = exclusive scan of eq
__global__ void __launch_bounds__(256)
dp_resolve_kernel(int n, const int* __restrict__ gt, const int* __restrict__ eq_flag, const int* __restrict__ eq_rank,
                  const DpState* __restrict__ S, int* __restrict__ sel) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  sel[i] = (gt[i] || (eq_flag[i] && eq_rank[i] < S->k_rem)) ? 1 : 0;
}

// keep = use_top64 ? sel : keep0   (in place into keep0); also a copy to scan
__global__ void __launch_bounds__(256)
dp_merge_keep_kernel(int n, int* __restrict__ keep0, const int* __restrict__ sel, const DpState* __restrict__ S,
                     int* __restrict__ scan_copy) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int k = S->use_top64 ? sel[i] : keep0[i];
  keep0[i] = k;
  scan_copy[i] = k;
}

__global__ void dp_total_kernel(const int* __restrict__ scan, const int* __restrict__ flag, int n, int* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *out = n > 0 ? scan[n - 1] + flag[n - 1] : 0;
}

// ---- Philox4x32-10 (counter-based) + Box-Muller ---------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}
__device__ __forceinline__ float u01(unsigned x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }   // (0,1)

__global__ void __launch_bounds__(256)
dp_scatter_kernel(int n, int cf, const float* __restrict__ means, const float* __restrict__ scales_raw,
                  const float* __restrict__ op_raw, const float* __restrict__ colors, const int* __restrict__ keep,
                  const int* __restrict__ keep_pos, const int* __restrict__ sel, const int* __restrict__ sel_pos,
                  const DpState* __restrict__ S, unsigned seed_lo, unsigned seed_hi, unsigned iter_lo, unsigned iter_hi,
                  float* __restrict__ o_means, float* __restrict__ o_scales, float* __restrict__ o_op,
                  float* __restrict__ o_colors) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n || !keep[i]) return;
  const float m0 = means[3 * (size_t)i], m1 = means[3 * (size_t)i + 1], m2 = means[3 * (size_t)i + 2];
  const float s0 = scales_raw[3 * (size_t)i], s1 = scales_raw[3 * (size_t)i + 1], s2 = scales_raw[3 * (size_t)i + 2];
  const float o = op_raw[i];
  {
    const size_t d = (size_t)keep_pos[i];
    o_means[3 * d] = m0; o_means[3 * d + 1] = m1; o_means[3 * d + 2] = m2;
    o_scales[3 * d] = s0; o_scales[3 * d + 1] = s1; o_scales[3 * d + 2] = s2;
    o_op[d] = o;
    for (int q = 0; q < cf; ++q) o_colors[d * cf + q] = colors[(size_t)i * cf + q];
  }
  if (sel[i]) {
    const size_t d = (size_t)S->n1 + sel_pos[i];
    const uint4 r = philox4x32_10(make_uint4((unsigned)i, iter_lo, iter_hi, 0x3D6A55u), make_uint2(seed_lo, seed_hi));
    // Box-Muller: two normals from (r.x, r.y), one from (r.z, r.w)
    const float ra = sqrtf(-2.0f * logf(u01(r.x))), rb = sqrtf(-2.0f * logf(u01(r.z)));
    float sn, cs, sn2, cs2;
    sincospif(2.0f * u01(r.y), &sn, &cs);
    sincospif(2.0f * u01(r.w), &sn2, &cs2);
    const float z0 = ra * cs, z1 = ra * sn, z2 = rb * cs2;
    (void)sn2;
    o_means[3 * d] = m0 + 0.25f * (softplusf_acc(s0) + 1e-3f) * z0;
    o_means[3 * d + 1] = m1 + 0.25f * (softplusf_acc(s1) + 1e-3f) * z1;
    o_means[3 * d + 2] = m2 + 0.25f * (softplusf_acc(s2) + 1e-3f) * z2;
    o_scales[3 * d] = s0; o_scales[3 * d + 1] = s1; o_scales[3 * d + 2] = s2;
    o_op[d] = o - 0.1f;
    for (int q = 0; q < cf; ++q) o_colors[d * cf + q] = colors[(size_t)i * cf + q];
  }
}

__global__ void dp_sum_kernel(const DpState* S, int* out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *out = S->n1 + S->add_n;
}

struct DpLayout { size_t state, keys, keep, sel, gt, eq, scanA, scanB, bs, total; };
static DpLayout dp_layout(int n) {
  const size_t nn = (size_t)(n > 0 ? n : 1);
  DpLayout L;
  size_t o = 0;
  L.state = o; o += align_up(sizeof(DpState));
  L.keys = o;  o += align_up(nn * 4);
  L.keep = o;  o += align_up(nn * 4);
  L.sel = o;   o += align_up(nn * 4);
  L.gt = o;    o += align_up(nn * 4);
  L.eq = o;    o += align_up(nn * 4);
  L.scanA = o; o += align_up(nn * 4);
  L.scanB = o; o += align_up(nn * 4);
  L.bs = o;    o += align_up((nn / 4096 + 2) * 4);
  L.total = o;
  return L;
}
size_t densify_workspace_bytes(int n) { return dp_layout(n).total; }

// radix select of the S->k_rem largest keys among `cand` (null = all) -> sel flags
static int dp_select(int n, const unsigned* keys, const int* cand, DpState* S, int* gt, int* eq, int* eq_rank, int* bs,
                     int* sel, cudaStream_t st) {
  const int blocks = (n + 255) / 256;
  const int hb = blocks < 592 ? blocks : 592;
  for (int shift = 24; shift >= 0; shift -= 8) {
    dp_hist_kernel<<<hb, 256, 0, st>>>(n, keys, cand, S, shift);
    B2S_LAUNCH_CHECK();
    dp_pick_kernel<<<1, 32, 0, st>>>(S, shift);
    B2S_LAUNCH_CHECK();
  }
  dp_mark_kernel<<<blocks, 256, 0, st>>>(n, keys, cand, S, gt, eq);
  B2S_LAUNCH_CHECK();
  B2S_CUDA_TRY(cudaMemcpyAsync(eq_rank, eq, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
  int rc = launch_scan_i32(eq_rank, n, bs, st);
  if (rc != B2S_OK) return rc;
  dp_resolve_kernel<<<blocks, 256, 0, st>>>(n, gt, eq, eq_rank, S, sel);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int launch_densify_prune(const float* means, const float* scales_raw, const float* op_raw, const float* colors, int n,
                         int color_floats, int max_gaussians, double ratio, float prune_opacity, unsigned long long seed,
                         unsigned long long iter, float* o_means, float* o_scales, float* o_op, float* o_colors,
                         int* n_new_dev, void* ws, cudaStream_t st) {
  const DpLayout L = dp_layout(n);
  char* w = (char*)ws;
  DpState* S = (DpState*)(w + L.state);
  unsigned* keys = (unsigned*)(w + L.keys);
  int *keep = (int*)(w + L.keep), *sel = (int*)(w + L.sel), *gt = (int*)(w + L.gt), *eq = (int*)(w + L.eq);
  int *scanA = (int*)(w + L.scanA), *scanB = (int*)(w + L.scanB), *bs = (int*)(w + L.bs);
  B2S_CUDA_TRY(cudaMemsetAsync(S, 0, sizeof(DpState), st));
  if (n <= 0) {
    B2S_CUDA_TRY(cudaMemsetAsync(n_new_dev, 0, 4, st));
    return B2S_OK;
  }
  const int blocks = (n + 255) / 256;
  dp_keys_kernel<<<blocks, 256, 0, st>>>(n, op_raw, prune_opacity, keys, keep, S);
  B2S_LAUNCH_CHECK();
  // prune: the threshold set, or the top 64 when it is too small
  dp_decide_kernel<<<1, 32, 0, st>>>(S, 0, n, max_gaussians, ratio, nullptr);
  B2S_LAUNCH_CHECK();
  int rc = dp_select(n, keys, nullptr, S, gt, eq, scanB, bs, sel, st);
  if (rc != B2S_OK) return rc;
  dp_merge_keep_kernel<<<blocks, 256, 0, st>>>(n, keep, sel, S, scanA);
  B2S_LAUNCH_CHECK();
  rc = launch_scan_i32(scanA, n, bs, st);            // scanA = position of each survivor
  if (rc != B2S_OK) return rc;
  dp_total_kernel<<<1, 32, 0, st>>>(scanA, keep, n, n_new_dev);
  B2S_LAUNCH_CHECK();
  // densify: the add_n most opaque survivors
  dp_decide_kernel<<<1, 32, 0, st>>>(S, 1, n, max_gaussians, ratio, n_new_dev);
  B2S_LAUNCH_CHECK();
  rc = dp_select(n, keys, keep, S, gt, eq, scanB, bs, sel, st);
  if (rc != B2S_OK) return rc;
  B2S_CUDA_TRY(cudaMemcpyAsync(scanB, sel, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
  rc = launch_scan_i32(scanB, n, bs, st);            // scanB = position of each clone
  if (rc != B2S_OK) return rc;
  dp_scatter_kernel<<<blocks, 256, 0, st>>>(n, color_floats, means, scales_raw, op_raw, colors, keep, scanA, sel, scanB, S,
                                            (unsigned)seed, (unsigned)(seed >> 32), (unsigned)iter, (unsigned)(iter >> 32),
                                            o_means, o_scales, o_op, o_colors);
  B2S_LAUNCH_CHECK();
  dp_sum_kernel<<<1, 32, 0, st>>>(S, n_new_dev);     // n_new = survivors + clones
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

}  // namespace b2s
