/* b2splat -- C-ABI of the B200-native Gaussian-splatting rasterizer (libb2splat.so).
 *
 * This header is the drop-in boundary for the render path of Kirkice/3DGaussian.
 * It replaces, for that path only:
 *   - the native entry   gr::render_gaussians          (reference include/gr/renderer.h:33-39,
 *                                                        src/renderer_dispatch.cpp:5-21)
 *   - the pybind module  gaussian_renderer             (reference src/bindings.cpp:27-100)
 *   - and it is what a torch.autograd.Function binds to stand in for
 *     render_gaussians_torch                           (reference python/torch_renderer.py:109-203)
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer owned by the caller unless the
 *     name ends in _host; nothing is allocated or freed inside the hot calls except by
 *     b2s_render_rgba8_host, which keeps a grow-only device cache in the ctx (the role of
 *     the reference's static DeviceBuffers, src/renderer.cu:287-349);
 *   - work is enqueued on the caller's CUDA stream (pass the raw cudaStream_t as void*);
 *     no implicit synchronisation unless documented;
 *   - return value 0 = success, negative = error; the message is available from
 *     b2s_last_error() (thread-local).  Nothing throws across the ABI.  There is no CPU
 *     fallback: params.force_cpu != 0 is rejected with B2S_ERR_UNSUPPORTED.
 *   - matrices are 4x4 row-major float, column-vector convention (p_cam = view * [m,1]),
 *     exactly as RenderParams (reference include/gr/gaussian_types.h:24-46).
 */
#ifndef B2SPLAT_H_
#define B2SPLAT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2S_OK 0
#define B2S_ERR_INVALID -1      /* bad argument */
#define B2S_ERR_CUDA -2         /* CUDA runtime error, see b2s_last_error */
#define B2S_ERR_UNSUPPORTED -3  /* e.g. force_cpu, unsupported tile size */
#define B2S_ERR_WORKSPACE -4    /* workspace / state buffer too small */
#define B2S_ERR_OVERFLOW -5     /* more (Gaussian,tile) pairs than max_pairs (only from syncing calls) */

#define B2S_TILE 16             /* pixels per tile edge */

/* blend modes (reference include/gr/gaussian_types.h:30-34 `enable_depth_sort`) */
#define B2S_MODE_WSUM 0         /* order-independent weighted sum (torch_renderer.py, renderer_cpu.cpp mode 0) */
#define B2S_MODE_SORTED 1       /* depth-sorted front-to-back "over" (renderer_cpu.cpp:125-217) */

/* arithmetic style of the per-Gaussian projection */
#define B2S_STYLE_TORCH 0       /* python/torch_renderer.py:57-78,147-150 (|s|, max(|z|,1e-6), divide by w_safe, op clamp_min 0) */
#define B2S_STYLE_NATIVE 1      /* src/renderer_cpu.cpp:166-200 (signed s, |z|+1e-6, multiply by 1/w, skip a<1e-5) */

/* activation flags: parameters are passed raw and activated inside the preprocess kernel
 * (python/fit_multiview_stub.py:268-275) */
#define B2S_ACT_SCALES_SOFTPLUS 1   /* scales = softplus(raw) + 1e-3 */
#define B2S_ACT_OPACITY_SIGMOID 2   /* opacities = sigmoid(raw) */
#define B2S_ACT_COLORS_SIGMOID 4    /* colors = sigmoid(raw), (N,3) colours only */

typedef struct b2s_ctx b2s_ctx;

/* Mirrors gr::RenderParams field for field, then the knobs the reference hard-codes. */
typedef struct b2s_params {
  int32_t width;
  int32_t height;
  float view[16];
  float proj[16];
  float background[3];
  int32_t enable_depth_sort; /* B2S_MODE_* */
  int32_t depth_slices;      /* accepted and ignored: sorting is exact, not sliced (src/renderer.cu:106-189) */
  int32_t force_cpu;         /* must be 0 */
  /* ---- extensions (the reference hard-codes these) ---- */
  int32_t style;             /* B2S_STYLE_* */
  float cutoff_sigma;        /* bbox half-size in sigmas; reference C++ uses 3 (renderer_cpu.cpp:96-97) */
  int32_t sh_coeffs;         /* colour floats per Gaussian / 3: 1 (RGB), 4 (reference "SH"), 9, 16 */
  int32_t sort_depth;        /* 1: radix-sort the full 64-bit tile|depth key; 0: tile bits only (stable) */
  int32_t act_flags;         /* B2S_ACT_* */
  int32_t exact_bbox;        /* 1: a Gaussian only touches pixels inside its bbox (renderer_cpu.cpp:202-203) */
  const float* background_dev; /* optional DEVICE pointer to 3 floats; when non-NULL it overrides background[] (the
                                * reference's callers pass the background as a device tensor,
                                * python/fit_multiview_stub.py:287: reading it back would cost a stream sync per view) */
  int32_t keep_depth;          /* 1: b2s_forward accumulates the depth plane into `state` even when out_depth is NULL
                                * (the fit loop's depth loss reads it from there) */
  int32_t reserved_;
} b2s_params;

/* ---- lifetime -------------------------------------------------------------------------- */
b2s_ctx* b2s_create(int device);
void b2s_destroy(b2s_ctx* ctx);
const char* b2s_last_error(void);
const char* b2s_version(void);

/* ---- sizes ------------------------------------------------------------------------------ */
/* bytes of the per-view state saved by b2s_forward for b2s_backward */
size_t b2s_state_bytes(int n, int width, int height, int64_t max_pairs);
/* bytes of transient scratch needed by any call below */
size_t b2s_workspace_bytes(int n, int width, int height, int64_t max_pairs);

/* Number of (Gaussian,tile) pairs this view produces.  Runs the projection + count and
 * SYNCHRONISES the stream to return the total on the host. */
int b2s_count_pairs(b2s_ctx* ctx, const b2s_params* p, const float* means, const float* scales,
                    const float* opacities, int n, int64_t* total_host, void* workspace,
                    size_t ws_bytes, void* stream);

/* ---- differentiable render (stands in for render_gaussians_torch) --------------------- */
/* out_rgb (H,W,3), out_alpha (H,W), out_depth (H,W) float32; alpha/depth may be NULL, and in the
 * weighted-sum mode out_rgb may be NULL too (the fit loop: the accumulators kept in `state` are all
 * b2s_fit_backward_blend needs).  colors: (N,3) or (N,sh_coeffs,3).  If the view needs more than max_pairs pairs the extra
 * ones are dropped and state's overflow counter is set (query with b2s_state_info). */
int b2s_forward(b2s_ctx* ctx, const b2s_params* p, const float* means, const float* scales,
                const float* colors, const float* opacities, int n, int64_t max_pairs,
                float* out_rgb, float* out_alpha, float* out_depth, void* state, size_t state_bytes,
                void* workspace, size_t ws_bytes, void* stream);

/* Gradients of sum(out_rgb*g_rgb + out_alpha*g_alpha + out_depth*g_depth) wrt the inputs.
 * g_alpha / g_depth may be NULL (treated as zeros).  accumulate != 0: grads are added to
 * the output buffers, else overwritten.  grad_scales is (N,3) with column 2 == 0. */
int b2s_backward(b2s_ctx* ctx, const b2s_params* p, const float* means, const float* scales,
                 const float* colors, const float* opacities, int n, int64_t max_pairs,
                 const float* g_rgb, const float* g_alpha, const float* g_depth, const void* state,
                 void* workspace, size_t ws_bytes, float* grad_means, float* grad_scales,
                 float* grad_colors, float* grad_opacities, int accumulate, void* stream);

/* ---- extension modes (named by north_star, ABSENT from the reference: SURVEY.md section 0) ---------------------
 * The reference has no rotation, no covariance projection and no differentiable sorted compositing
 * (python/torch_renderer.py:109-121 takes no `rotations`; include/gr/gaussian_types.h:8-22 has no such field;
 * src/renderer_cpu.cpp:125-217 is not differentiable), so nothing on the reference side binds these two entry points:
 * they extend b2s_forward / b2s_backward for callers that pass the new keyword arguments of the Python drop-in
 * (render_gaussians_torch(..., rotations=, blend=)).  Pinned by oracle/ext_oracle.py only.
 *   rotations: (N,4) quaternions (w,x,y,z), normalised inside, or NULL for the reference's axis-aligned sigmas.
 *              With rotations the footprint is the EWA projection  cov2 = J W R diag(s^2) R^T W^T J^T + ewa_dilation*I
 *              (J = exact Jacobian of the reference's pixel projection, W = view[:3,:3]; all three scale columns used).
 *   blend:     B2S_BLEND_WSUM  out = (bg + sum w c)/(1 + sum w), alpha, depth as b2s_forward;
 *              B2S_BLEND_OVER  front-to-back by camera z with the rule of src/renderer_cpu.cpp:196-215 (exact
 *                              cutoff_sigma pixel bbox, a < 1e-5 skipped, contrib = T a, a capped at 0.999999),
 *                              rgb = C + T bg, alpha = 1 - T, depth = sum T a z (expected depth).
 * params->style / enable_depth_sort / sort_depth / exact_bbox are ignored (torch-style inputs; order by `blend`).
 * state / workspace sizes as for b2s_forward.  grad_scales is (N,3) (column 2 is zero without rotations),
 * grad_rotations (N,4) or NULL when rotations is NULL.  Gradients are overwritten, not accumulated. */
#define B2S_BLEND_WSUM 0
#define B2S_BLEND_OVER 1
int b2s_forward_ext(b2s_ctx* ctx, const b2s_params* p, const float* means, const float* scales, const float* rotations,
                    const float* colors, const float* opacities, int n, int64_t max_pairs, int blend, float ewa_dilation,
                    float* out_rgb, float* out_alpha, float* out_depth, void* state, size_t state_bytes,
                    void* workspace, size_t ws_bytes, void* stream);
int b2s_backward_ext(b2s_ctx* ctx, const b2s_params* p, const float* means, const float* scales, const float* rotations,
                     const float* colors, const float* opacities, int n, int64_t max_pairs, int blend, float ewa_dilation,
                     const float* g_rgb, const float* g_alpha, const float* g_depth, const void* state, void* workspace,
                     size_t ws_bytes, float* grad_means, float* grad_scales, float* grad_rotations, float* grad_colors,
                     float* grad_opacities, void* stream);

/* ---- multi-view backward (the fit loop) --------------------------------------------------
 * b2s_backward = b2s_backward_blend (per view) + b2s_backward_params (chain rule).  The fit loop
 * calls the first once per view into slot v of a (num_views, n, 12) float buffer and the second
 * ONCE per iteration: every Gaussian's gradients are accumulated over all views in registers
 * and written once, instead of a read-modify-write of the whole gradient buffer per view. */
size_t b2s_view_block_bytes(void);
/* Converts num_views parameter blocks into the kernels' per-view constant blocks
 * (out_host: num_views * b2s_view_block_bytes() bytes of HOST memory; upload it once). */
int b2s_pack_views(const b2s_params* params, int num_views, void* out_host);
/* Blend backward of one view: gacc_out (n,12) float = per-Gaussian partial sums (overwritten), opaque to the caller:
 * rows {dR, dG, dB, Syy | S, Sx, Sxx, Sy | dZ, colour clamp mask, -, -} consumed by b2s_backward_params. */
int b2s_backward_blend(b2s_ctx* ctx, const b2s_params* p, int n, int64_t max_pairs, const float* g_rgb,
                       const float* g_alpha, const float* g_depth, const void* state, void* workspace,
                       size_t ws_bytes, float* gacc_out, void* stream);
/* b2s_fit_loss + b2s_backward_blend in one call (python/fit_multiview_stub.py:292-303 and its backward): the
 * per-view loss  mean|rgb-tgt| + w_sil*mean|alpha-mask| + w_depth*mean|depth/(max depth + 1e-6) - depth_gt|  is
 * evaluated from the accumulators b2s_forward left in `state` and its image gradients, scaled by `scale` (1/V), go
 * straight into the blend backward; no rgb / alpha / depth / gradient images are read or written.  mask and
 * depth_gt may be NULL (term absent); the depth term needs a forward run with params.keep_depth = 1 (and a cutoff of
 * 7 sigma).  Adds scale*loss to *loss_accum. */
int b2s_fit_backward_blend(b2s_ctx* ctx, const b2s_params* p, int n, int64_t max_pairs, const float* tgt,
                           const float* mask, const float* depth_gt, float w_sil, float w_depth, float scale,
                           float* loss_accum, const void* state,
                           const void* prepared_view /* NULL unless the forward was b2s_forward_prepared */,
                           void* workspace, size_t ws_bytes, float* gacc_out, void* stream);
/* The same with the target and the mask as the 8-bit image bytes they are on disk (the reference converts them on the
 * host, np.asarray(img, float32) / 255, python/fit_multiview_stub.py:16-23): the loss kernel reads the bytes and divides
 * by 255 itself -- no float32 copy of the target exists on the device and a host-fed fit moves 1 byte per value over
 * PCIe.  mask_u8 may be NULL; depth_gt stays float32. */
int b2s_fit_backward_blend_u8(b2s_ctx* ctx, const b2s_params* p, int n, int64_t max_pairs, const uint8_t* tgt_u8,
                              const uint8_t* mask_u8, const float* depth_gt, float w_sil, float w_depth, float scale,
                              float* loss_accum, const void* state, const void* prepared_view, void* workspace,
                              size_t ws_bytes, float* gacc_out, void* stream);

/* Batched per-Gaussian stage of the fit loop: projection, sigma, SH colour, tile rect and tile mask of EVERY
 * local view in one launch -- the parameters (192 B of SH coefficients per Gaussian at degree 3) are read once
 * per iteration instead of once per view.  views_dev = device copy of the b2s_pack_views block; `prepared`
 * receives num_views blocks of b2s_prepared_view_bytes(n) bytes.  b2s_forward_prepared then renders view v
 * from block v (weighted-sum mode, sort_depth = 0 only) exactly as b2s_forward would from the parameters. */
size_t b2s_prepared_view_bytes(int n);
int b2s_preprocess_views(b2s_ctx* ctx, const void* views_dev, int num_views, int sh_coeffs, const float* means,
                         const float* scales, const float* colors, const float* opacities, int n, void* prepared,
                         void* stream);
int b2s_forward_prepared(b2s_ctx* ctx, const b2s_params* p, const void* prepared_view, int n, int64_t max_pairs,
                         float* out_rgb, float* out_alpha, float* out_depth, void* state, size_t state_bytes,
                         void* workspace, size_t ws_bytes, void* stream);
/* Chain rule over all views: views_dev = device copy of the b2s_pack_views block. */
int b2s_backward_params(b2s_ctx* ctx, const void* views_dev, int num_views, int sh_coeffs, const float* means,
                        const float* scales, const float* colors, const float* opacities, int n,
                        const float* gacc_all, float* grad_means, float* grad_scales, float* grad_colors,
                        float* grad_opacities, int accumulate, void* stream);

/* The same for the Gaussians [first, first + count) only (all pointers are the bases of the full arrays): the
 * multi-GPU fit folds the views chunk by chunk, so that chunk k's gradients cross NVLink (all-reduce) and take their
 * Adam step while chunk k+1 is still being computed. */
int b2s_backward_params_range(b2s_ctx* ctx, const void* views_dev, int num_views, int sh_coeffs, const float* means,
                              const float* scales, const float* colors, const float* opacities, int n, int first,
                              int count, const float* gacc_all, float* grad_means, float* grad_scales,
                              float* grad_colors, float* grad_opacities, int accumulate, void* stream);

/* info_host[0] = pairs needed, [1] = pairs kept, [2] = overflow flag.  Synchronises. */
int b2s_state_info(b2s_ctx* ctx, const void* state, int n, int width, int height, int64_t max_pairs,
                   int64_t* info_host, void* stream);

/* ---- RGBA8 frame (stands in for gr::render_gaussians) --------------------------------- */
/* device-resident inputs, device output (H,W,4) uint8 */
int b2s_render_rgba8(b2s_ctx* ctx, const b2s_params* p, const float* means, const float* scales,
                     const float* colors, const float* opacities, int n, int64_t max_pairs,
                     uint8_t* out_rgba, void* workspace, size_t ws_bytes, void* stream);
/* HOST pointers in and out, same argument meaning as gr::render_gaussians; synchronous. */
int b2s_render_rgba8_host(b2s_ctx* ctx, const b2s_params* p, const float* means_host,
                          const float* scales_host, const float* colors_host,
                          const float* opacities_host, int n, uint8_t* out_rgba_host);

/* ---- binning dump (bit-exact tests) ------------------------------------------------------ */
/* Runs projection -> count -> scan -> key emit -> radix sort -> tile ranges and copies the
 * intermediate results into caller buffers (any may be NULL):
 *   px,py,sx,sy,zabs: float[n]; bbox: int32[4n]; cnt: int32[n];
 *   keys_unsorted/keys_sorted: uint64[max_pairs]; vals_unsorted/vals_sorted: int32[max_pairs];
 *   ranges: int32[2*tiles]; total: int64[1] (device). */
int b2s_dump_bins(b2s_ctx* ctx, const b2s_params* p, const float* means, const float* scales,
                  const float* opacities, int n, int64_t max_pairs, float* px, float* py, float* sx,
                  float* sy, float* zabs, int32_t* bbox, int32_t* cnt, uint64_t* keys_unsorted,
                  int32_t* vals_unsorted, uint64_t* keys_sorted, int32_t* vals_sorted, int32_t* ranges,
                  int64_t* total, void* workspace, size_t ws_bytes, void* stream);

/* Stand-alone stable LSD radix sort of (uint64 key, int32 value) pairs on key bits
 * [begin_bit, end_bit); result in keys_out/vals_out.  tmp: b2s_sort_tmp_bytes(m). */
size_t b2s_sort_tmp_bytes(int64_t m);
int b2s_sort_pairs(b2s_ctx* ctx, const uint64_t* keys_in, const int32_t* vals_in, uint64_t* keys_out,
                   int32_t* vals_out, int64_t m, int begin_bit, int end_bit, void* tmp, size_t tmp_bytes,
                   void* stream);

/* ---- fit-loop kernels (python/fit_multiview_stub.py:292-311) --------------------------- */
/* Per-view loss  mean|rgb-tgt| + w_sil*mean|alpha-mask|  and its gradient wrt rgb/alpha,
 * scaled by `scale` (1/V).  mask/g_alpha may be NULL.  Adds the loss value to *loss_accum. */
int b2s_fit_loss(b2s_ctx* ctx, const float* rgb, const float* alpha, const float* tgt,
                 const float* mask, int width, int height, float w_sil, float scale,
                 float* g_rgb, float* g_alpha, float* loss_accum, void* stream);

/* Target ingestion: dst[i] = src[i] / 255 for `count` 8-bit values (np.asarray(img, float32) / 255.0,
 * python/fit_multiview_stub.py:16-23), so host-fed targets cross PCIe as bytes. */
int b2s_u8_to_f32(b2s_ctx* ctx, const uint8_t* src, float* dst, int64_t count, void* stream);

/* torch.optim.Adam step (defaults beta=(0.9,0.999), eps=1e-8, no weight decay) over a flat
 * parameter buffer; step is 1-based.  Optional fused regulariser gradients of
 *   reg_opacity*mean(sigmoid(op_raw)) + reg_scale*mean(softplus(scales_raw)+1e-3)
 * (fit_multiview_stub.py:307) for the element ranges [scales_begin,scales_end),
 * [opac_begin,opac_end) of the flat buffer (pass empty ranges to disable). */
int b2s_adam_step(b2s_ctx* ctx, float* params, const float* grads, float* m, float* v, int64_t count,
                  int step, float lr, float beta1, float beta2, float eps, int64_t scales_begin,
                  int64_t scales_end, float reg_scale, int64_t opac_begin, int64_t opac_end,
                  float reg_opacity, void* stream);

/* The same step, guarded on the device: if *skip_flag != 0 (a float the caller's kernels / collective left on the
 * device, e.g. the summed pair-buffer overflow count of the iteration's views) NOTHING is updated and
 * *skipped_count is incremented instead -- an iteration whose views overflowed must not reach the parameters, and the
 * host need not synchronise every step to guarantee it (FitDriver polls skipped_count now and then, re-plans the
 * buffers and repeats the skipped iterations).  skip_flag / skipped_count may be NULL (= b2s_adam_step). */
int b2s_adam_step_guarded(b2s_ctx* ctx, float* params, const float* grads, float* m, float* v, int64_t count,
                          int step, float lr, float beta1, float beta2, float eps, int64_t scales_begin,
                          int64_t scales_end, float reg_scale, int64_t opac_begin, int64_t opac_end,
                          float reg_opacity, const float* skip_flag, int* skipped_count, void* stream);

/* Multi-GPU form of the guarded Adam step: gradient reduce-scatter + Adam + parameter all-gather in ONE kernel over
 * NVLink multicast (NVLS multimem.ld_reduce / multimem.st).  The reference has no multi-GPU code;
 * on one GPU this is b2s_adam_step_guarded = torch.optim.Adam of python/fit_multiview_stub.py:262,311.
 *   params_mc / grads_mc: MULTICAST addresses of this slice in the symmetric parameter / gradient buffers of all `world`
 *   ranks (every rank passes the same offsets); params_local, m, v: this rank's own memory for the same slice.
 * Rank r updates the r-th share of the slice (ceil(count/world) rounded up to 4 floats) from the switch-reduced gradient
 * and multicasts the new parameters; moments of a share are kept by its owner only.  scales_/opac_ ranges are relative to
 * the slice, as in b2s_adam_step.  The caller separates the chain rule that wrote the gradients, these launches and the
 * next reader of the parameters by cross-rank barriers.  b2s_reduce_tail_multimem: the first `count` (<= 64) floats at
 * tail_mc summed over the ranks into tail_out (local) -- the iteration's loss and the overflow count that guards the step. */
int b2s_adam_step_multimem(b2s_ctx* ctx, float* params_mc, const float* grads_mc, const float* params_local, float* m,
                           float* v, int64_t count, int rank, int world, int step, float lr, float beta1, float beta2,
                           float eps, int64_t scales_begin, int64_t scales_end, float reg_scale, int64_t opac_begin,
                           int64_t opac_end, float reg_opacity, const float* skip_flag, int* skipped_count, void* stream);
int b2s_reduce_tail_multimem(b2s_ctx* ctx, const float* tail_mc, float* tail_out, int count, void* stream);
/* Host helper: the share [lo, hi) of a slice of `count` floats that rank `rank` of `world` updates in b2s_adam_step_multimem
 * (FitDriver gathers the owners' Adam moments with it; no CUDA call). */
int b2s_multimem_share(int64_t count, int rank, int world, int64_t* lo, int64_t* hi);

/* Densify / prune (python/fit_multiview_stub.py:140-197): keep sigmoid(op_raw) > prune_opacity (or the
 * 64 most opaque if fewer survive), order preserved; then append min(max_gaussians - n1, int(n1 *
 * densify_ratio)) clones of the most opaque survivors with mean + 0.25*scale*N(0,1), op_raw - 0.1.
 * The N(0,1) draws are Philox4x32-10(seed, iteration, source index): identical on every rank.
 * Outputs must hold max(max_gaussians, n) Gaussians; colours have color_floats floats per Gaussian.
 * The caller resets its Adam state, as the reference does (:319-325).  SYNCHRONISES (returns the new
 * count in *n_new_host). */
size_t b2s_densify_workspace_bytes(int n);
int b2s_densify_prune(b2s_ctx* ctx, const float* means, const float* scales_raw, const float* opacities_raw,
                      const float* colors, int n, int color_floats, int max_gaussians, double densify_ratio,
                      float prune_opacity, uint64_t seed, uint64_t iteration, float* out_means,
                      float* out_scales_raw, float* out_opacities_raw, float* out_colors, int* n_new_host,
                      void* workspace, size_t ws_bytes, void* stream);

/* ---- forward tickets: cheap, exact overflow detection ---------------------------------------
 * Every b2s_forward / b2s_forward_prepared / b2s_render_rgba8 call issued through `ctx` gets a ticket (a running
 * number, returned by b2s_last_ticket right after the call).  The binning kernels of that call mirror their pair
 * counters into pinned host memory and an event is recorded right behind them, so b2s_ticket_info waits for THAT
 * point of the stream only -- the blend kernels queued behind it keep running -- and returns
 * info_host[0] = pairs needed, [1] = pairs kept, [2] = overflow flag.  The drop-in renderer sizes its pair buffers
 * from a cached capacity and re-renders the rare view that overflowed, instead of a count pass + full sync per call.
 * The ring holds the last 256 tickets of a ctx. */
int64_t b2s_last_ticket(b2s_ctx* ctx);
int b2s_ticket_info(b2s_ctx* ctx, int64_t ticket, int64_t* info_host);

/* ---- instrumentation -------------------------------------------------------------------- */
/* number of kernels this library has launched in this process (all contexts) */
int64_t b2s_launch_count(void);
/* launches of the four weighted-sum blend kernel families so far: out[0] = forward tcgen05, [1] = forward other
 * (mma.sync / FP32), [2] = backward tcgen05, [3] = backward other.  Tests use it to prove which path ran. */
void b2s_path_counts(int64_t out[4]);
/* SM count the persistent grids are sized from (cudaDevAttrMultiProcessorCount of the current device) */
int b2s_sm_count(void);
/* Per-stage CUDA-event timing.  While enabled, every stage launched through `ctx` is
 * bracketed by events on the caller's stream.  b2s_timing_read synchronises on the last
 * event, adds up elapsed milliseconds and launch counts per stage (arrays of
 * b2s_num_stages() entries) and clears the log. */
int b2s_num_stages(void);
const char* b2s_stage_name(int stage);
int b2s_timing_enable(b2s_ctx* ctx, int on);
int b2s_timing_read(b2s_ctx* ctx, float* ms_per_stage, int64_t* spans_per_stage);

#ifdef __cplusplus
}
#endif
#endif /* B2SPLAT_H_ */
