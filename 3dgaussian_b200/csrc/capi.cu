// extern "C" entry points of libb2splat.so -- see include/b2splat.h for the contract.
// Replaces the reference's dispatch + binding layer (src/renderer_dispatch.cpp:5-21,
// src/bindings.cpp:27-100) and the host wrapper of src/renderer.cu:272-408.
#include <math.h>
#include <stdarg.h>
#include <atomic>
#include <mutex>
#include <vector>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include "common.cuh"

constexpr int B2S_TICKET_RING = 256;

struct b2s_ctx {
  int device;
  // forward tickets (b2s_ticket_info): pair counters mirrored into pinned host memory + an event behind the binning
  b2s::Counters* probe = nullptr;                 // B2S_TICKET_RING slots, cudaHostAlloc (mapped)
  cudaEvent_t probe_ev[B2S_TICKET_RING] = {};
  long long tickets = 0;                          // tickets issued so far; the last one is tickets - 1
  // grow-only cache used by b2s_render_rgba8_host only
  void* host_dev = nullptr;
  size_t host_dev_bytes = 0;
  // optional per-stage event timing
  bool timing = false;
  struct Span { int stage; cudaEvent_t a, b; };
  std::vector<Span> spans;
  std::vector<cudaEvent_t> pool;
};

namespace b2s {

static thread_local char g_err[512] = "";

struct StageTimer {   // RAII: records an event pair around one stage when ctx->timing is on
  b2s_ctx* c;
  cudaStream_t st;
  cudaEvent_t b = nullptr;
  StageTimer(b2s_ctx* ctx, int stage, cudaStream_t s) : c(ctx), st(s) {
    if (c == nullptr || !c->timing) { c = nullptr; return; }
    cudaEvent_t e[2];
    for (int i = 0; i < 2; ++i) {
      if (!c->pool.empty()) { e[i] = c->pool.back(); c->pool.pop_back(); }
      else if (cudaEventCreate(&e[i]) != cudaSuccess) { c = nullptr; return; }
    }
    cudaEventRecord(e[0], st);
    b = e[1];
    c->spans.push_back({stage, e[0], e[1]});
  }
  ~StageTimer() { if (c != nullptr) cudaEventRecord(b, st); }
};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
static std::atomic<long long> g_paths[PATH_COUNT];
void count_path(int which) { if (which >= 0 && which < PATH_COUNT) g_paths[which].fetch_add(1, std::memory_order_relaxed); }

constexpr int MAX_DEVICES = 64;
static int current_device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) dev = MAX_DEVICES - 1;
  return dev;
}
int sm_count() {
  static std::atomic<int> cache[MAX_DEVICES];
  const int dev = current_device_slot();
  int v = cache[dev].load(std::memory_order_relaxed);
  if (v <= 0) {
    int real = 0;
    cudaGetDevice(&real);
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, real) != cudaSuccess || v <= 0) v = 1;
    cache[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}
std::once_flag& device_once_flag(int slot) {
  static std::once_flag flags[ONCE_SLOTS][MAX_DEVICES];
  return flags[slot][current_device_slot()];
}

// Default: the binning kernels write the mirror straight into mapped pinned memory (no extra stream operation);
// B2S_TICKET_MODE=copy: a 32-byte stream-ordered device-to-host copy of the counters behind the binning kernels
static bool ticket_mapped() {
  static const bool m = [] { const char* e = getenv("B2S_TICKET_MODE"); return !(e != nullptr && e[0] == 'c'); }();
  return m;
}
// Next ticket of the ctx: the slot of the pinned ring the binning kernels mirror their counters into (nullptr if
// the ring could not be allocated: tickets then report an error instead of data).
static Counters* ticket_begin(b2s_ctx* ctx) {
  if (ctx == nullptr) return nullptr;
  static const bool off = [] { const char* e = getenv("B2S_NO_TICKETS"); return e != nullptr && e[0] == '1'; }();
  if (off) return nullptr;
  if (ctx->probe == nullptr) {
    void* h = nullptr;
    if (cudaHostAlloc(&h, sizeof(Counters) * B2S_TICKET_RING, ticket_mapped() ? cudaHostAllocMapped : cudaHostAllocDefault) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    memset(h, 0, sizeof(Counters) * B2S_TICKET_RING);
    ctx->probe = (Counters*)h;
  }
  const int slot = (int)(ctx->tickets % B2S_TICKET_RING);
  ctx->tickets += 1;
  return ctx->probe + slot;
}
// records the event behind the kernels that wrote the mirror of the ticket just begun
static void ticket_mark(b2s_ctx* ctx, const Counters* counters_dev, cudaStream_t st) {
  if (ctx == nullptr || ctx->probe == nullptr || ctx->tickets <= 0) return;
  const int slot = (int)((ctx->tickets - 1) % B2S_TICKET_RING);
  if (!ticket_mapped() && cudaMemcpyAsync(ctx->probe + slot, counters_dev, sizeof(Counters), cudaMemcpyDeviceToHost, st) != cudaSuccess) {
    cudaGetLastError();
    return;
  }
  static const bool noev = [] { const char* e = getenv("B2S_TICKET_NOEVENT"); return e != nullptr && e[0] == '1'; }();
  if (noev) return;
  if (ctx->probe_ev[slot] == nullptr && cudaEventCreateWithFlags(&ctx->probe_ev[slot], cudaEventDisableTiming) != cudaSuccess) {
    ctx->probe_ev[slot] = nullptr;
    cudaGetLastError();
    return;
  }
  cudaEventRecord(ctx->probe_ev[slot], st);
}

// inverse of a row-major 4x4 (Gauss-Jordan, double) -> camera centre inv(V)[:3,3]
static bool camera_centre(const float* v, float* cam) {
  double a[4][8];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      a[r][c] = v[4 * r + c];
      a[r][4 + c] = (r == c) ? 1.0 : 0.0;
    }
  for (int col = 0; col < 4; ++col) {
    int piv = col;
    for (int r = col + 1; r < 4; ++r)
      if (fabs(a[r][col]) > fabs(a[piv][col])) piv = r;
    if (fabs(a[piv][col]) < 1e-30) return false;
    if (piv != col)
      for (int c = 0; c < 8; ++c) { const double t = a[col][c]; a[col][c] = a[piv][c]; a[piv][c] = t; }
    const double d = 1.0 / a[col][col];
    for (int c = 0; c < 8; ++c) a[col][c] *= d;
    for (int r = 0; r < 4; ++r)
      if (r != col) {
        const double f = a[r][col];
        if (f != 0.0)
          for (int c = 0; c < 8; ++c) a[r][c] -= f * a[col][c];
      }
  }
  for (int r = 0; r < 3; ++r) cam[r] = (float)a[r][7];
  return true;
}

static int make_view(const b2s_params* p, ViewParams* vp) {
  if (p == nullptr) { set_error("params is NULL"); return B2S_ERR_INVALID; }
  if (p->force_cpu != 0) { set_error("force_cpu is not supported: this library has no CPU path"); return B2S_ERR_UNSUPPORTED; }
  if (p->width <= 0 || p->height <= 0 || p->width > 32767 || p->height > 32767) {
    set_error("bad image size %dx%d", p->width, p->height);
    return B2S_ERR_INVALID;
  }
  if (!(p->cutoff_sigma > 0.0f)) { set_error("cutoff_sigma must be > 0"); return B2S_ERR_INVALID; }
  memcpy(vp->view, p->view, sizeof(float) * 16);
  memcpy(vp->proj, p->proj, sizeof(float) * 16);
  memcpy(vp->bg, p->background, sizeof(float) * 3);
  vp->bg_dev = p->background_dev;
  vp->keep_depth = p->keep_depth ? 1 : 0;
  vp->cam[0] = vp->cam[1] = vp->cam[2] = 0.0f;
  if (p->sh_coeffs > 1 && !camera_centre(p->view, vp->cam)) { set_error("view matrix is singular"); return B2S_ERR_INVALID; }
  vp->k = p->cutoff_sigma;
  vp->width = p->width;
  vp->height = p->height;
  vp->wm1 = (float)(p->width - 1);
  vp->hm1 = (float)(p->height - 1);
  vp->wf = (float)p->width;
  vp->hf = (float)p->height;
  vp->fx = fabsf(p->proj[0]);
  vp->fy = fabsf(p->proj[5]);
  vp->tiles_x = (p->width + TILE - 1) / TILE;
  vp->tiles_y = (p->height + TILE - 1) / TILE;
  vp->n_tiles = vp->tiles_x * vp->tiles_y;
  {
    static const int forced = [] { const char* e = getenv("B2S_UNIT_SIZE"); const int v = e ? atoi(e) : 0;
                                   return (v >= SEG_MIN && (v & (v - 1)) == 0) ? v : 0; }();   // development knob
    vp->seg = forced ? forced : unit_size(vp->n_tiles);
  }
  vp->style = p->style;
  vp->sh = p->sh_coeffs > 0 ? p->sh_coeffs : 1;
  vp->act = p->act_flags;
  vp->mode = p->enable_depth_sort ? B2S_MODE_SORTED : B2S_MODE_WSUM;
  vp->exact_bbox = (p->exact_bbox || vp->mode == B2S_MODE_SORTED) ? 1 : 0;
  return B2S_OK;
}

static int tile_bits(int n_tiles) {
  int b = 1;
  while ((1 << b) < n_tiles) ++b;
  return b;
}

struct Bufs {  // resolved pointers into state / workspace
  Counters* counters;
  float4* rec;
  int2* ranges;
  int* vals;
  float* acc;
  uint8_t* cmask;
  uint2* rect;
  unsigned long long* tmask;
  uint32_t* dbits;
  int* cnt;
  long long* bsum;
  unsigned long long *keysA, *keysB;
  int* valsB;
  int *hist, *hsum;
  float* gacc;
  int* unit_start;
  int2* units;
  int4* udesc;
  float* partial;
  float* gbuf;
  int* cs_table;
  int* cs_total;
  int *oorder, *oslab, *osmall, *ne_list;
  int2* rest_list;
  int64_t unit_cap;
};

static Bufs resolve(void* state, void* ws, int n, int w, int h, int64_t mp) {
  Bufs b;
  memset(&b, 0, sizeof(b));
  b.unit_cap = max_units(w, h, mp);
  if (state != nullptr) {
    const StateLayout S = state_layout(n, w, h, mp);
    char* s = (char*)state;
    b.counters = (Counters*)(s + S.counters);
    b.rec = (float4*)(s + S.rec);
    b.ranges = (int2*)(s + S.ranges);
    b.vals = (int*)(s + S.vals);
    b.acc = (float*)(s + S.acc);
    b.unit_start = (int*)(s + S.unit_start);
    b.units = (int2*)(s + S.units);
    b.udesc = (int4*)(s + S.udesc);
    b.cmask = (uint8_t*)(s + S.cmask);
  }
  if (ws != nullptr) {
    const WorkLayout L = work_layout(n, w, h, mp);
    char* q = (char*)ws;
    b.rect = (uint2*)(q + L.rect);
    b.tmask = (unsigned long long*)(q + L.tmask);
    b.dbits = (uint32_t*)(q + L.dbits);
    b.cnt = (int*)(q + L.cnt);
    b.bsum = (long long*)(q + L.bsum);
    b.keysA = (unsigned long long*)(q + L.keysA);
    b.keysB = (unsigned long long*)(q + L.keysB);
    b.valsB = (int*)(q + L.valsB);
    b.hist = (int*)(q + L.hist);
    b.hsum = (int*)(q + L.hsum);
    b.gacc = (float*)(q + L.gacc);
    b.partial = (float*)(q + L.partial);
    b.gbuf = (float*)(q + L.gbuf);
    b.cs_table = (int*)(q + L.cs_table);
    b.cs_total = (int*)(q + L.cs_total);
    b.oorder = (int*)(q + L.oorder);
    b.oslab = (int*)(q + L.oslab);
    b.osmall = (int*)(q + L.osmall);
    b.ne_list = (int*)(q + L.ne_list);
    b.rest_list = (int2*)(q + L.rest_list);
  }
  return b;
}

// Points the per-view inputs of the binning / blend stages at one view block of b2s_preprocess_views.
static void use_prepared(Bufs& b, const void* prepared_view, int n) {
  const PreparedLayout L = prepared_layout(n);
  char* q = (char*)const_cast<void*>(prepared_view);
  b.rec = (float4*)(q + L.rec);
  b.cmask = (uint8_t*)(q + L.cmask);
  b.rect = (uint2*)(q + L.rect);
  b.tmask = (unsigned long long*)(q + L.tmask);
}

// projection -> count -> scan -> emit -> sort -> ranges.  On return the sorted Gaussian ids are
// in B.vals (state) and the sorted keys in *keys_sorted.
static int run_binning(b2s_ctx* ctx, const ViewParams& vp, const b2s_params* p, const float* means, const float* scales,
                       const float* colors, const float* opac, int n, int64_t max_pairs, const Bufs& B,
                       float* dbg, int* dbg_bbox, unsigned long long** keys_sorted,
                       unsigned long long* keys_unsorted_copy, int* vals_unsorted_copy, cudaStream_t st,
                       bool pre_done = false /* the caller has already written rec / rect / tmask / dbits / cnt / bsum */) {
  int rc = B2S_OK;
  Counters* mirror = ticket_begin(ctx);
  if (!ticket_mapped()) mirror = nullptr;       // the counters travel by a stream-ordered copy instead (ticket_mark)
  if (pre_done) {
  } else if (means != nullptr) {    // means == NULL: B already points at a view block of b2s_preprocess_views
    StageTimer t(ctx, ST_PREPROCESS, st);
    rc = launch_preprocess(vp, means, scales, colors, opac, n, B.rec, B.cmask, B.rect, B.tmask, B.dbits, B.cnt, B.bsum, dbg, dbg_bbox, st);
  } else if (!counting_sort_fits(vp.n_tiles)) {
    set_error("prepared views feed the counting-sort path only (at most %d tiles)", (int)(CS_MAX_SMEM / 8));
    return B2S_ERR_UNSUPPORTED;
  }
  if (rc != B2S_OK) return rc;
  const int begin_bit = p->sort_depth ? 0 : 32;
  const int end_bit = 32 + tile_bits(vp.n_tiles);
  const int passes = sort_passes(begin_bit, end_bit);
  if (counting_sort_fits(vp.n_tiles)) {
    // group by tile with one counting pass + one scatter pass (the weighted sum needs no order inside a tile).  The
    // depth order, when asked for, comes from walking the Gaussians in their global depth order and sorting the inside
    // of the small (block, tile) groups (segsort.cu) instead of a 6-pass radix sort of all the pairs.
    const int* order = nullptr;
    const int* slab_start = nullptr;
    if (p->sort_depth) {
      if (means == nullptr) { set_error("the depth order needs the per-view preprocess (depth bits)"); return B2S_ERR_UNSUPPORTED; }
      StageTimer t(ctx, ST_SORT, st);
      uint32_t* splitters = (uint32_t*)B.osmall;
      int *count = B.osmall + 1024, *sstart = B.osmall + 2048, *cursor = B.osmall + 3072;
      rc = launch_depth_slabs(B.dbits, n, counting_sort_blocks(n), splitters, count, sstart, cursor, B.oslab, B.oorder, st);
      if (rc != B2S_OK) return rc;
      order = B.oorder;
      slab_start = sstart;
    }
    {
      StageTimer t(ctx, ST_BIN, st);
      rc = launch_counting_sort(vp, n, max_pairs, B.rect, B.tmask, order, slab_start, B.cs_table, B.cs_total, B.ranges, B.counters, mirror, B.unit_cap,
                                B.unit_start, B.units, B.udesc, B.ne_list, B.vals, 0, st);   // also writes the unit descriptor table
    }
    if (rc != B2S_OK) return rc;
    ticket_mark(ctx, B.counters, st);
    if (keys_unsorted_copy != nullptr || vals_unsorted_copy != nullptr) {   // dump hook only: the emit order
      Counters* scratch = (Counters*)B.hist;
      rc = launch_bin(vp, n, max_pairs, B.rect, B.tmask, B.dbits, B.cnt, B.bsum, B.keysA, B.valsB, scratch, nullptr, st);
      if (rc != B2S_OK) return rc;
      if (keys_unsorted_copy != nullptr)
        B2S_CUDA_TRY(cudaMemcpyAsync(keys_unsorted_copy, B.keysA, (size_t)max_pairs * 8, cudaMemcpyDeviceToDevice, st));
      if (vals_unsorted_copy != nullptr)
        B2S_CUDA_TRY(cudaMemcpyAsync(vals_unsorted_copy, B.valsB, (size_t)max_pairs * 4, cudaMemcpyDeviceToDevice, st));
    }
    {
      StageTimer t(ctx, ST_SORT, st);
      rc = launch_counting_sort(vp, n, max_pairs, B.rect, B.tmask, order, slab_start, B.cs_table, B.cs_total, B.ranges, B.counters, nullptr, B.unit_cap,
                                B.unit_start, B.units, B.udesc, B.ne_list, B.vals, 1, st);
      if (rc == B2S_OK && p->sort_depth)
        rc = launch_group_sort(vp, B.cs_table, B.cs_total, B.ranges, B.ne_list, counting_sort_blocks(n), B.dbits, B.counters, B.vals, B.keysA,
                               keys_sorted != nullptr ? B.keysB : nullptr, st);
    }
    if (rc != B2S_OK) return rc;
    if (keys_sorted != nullptr) *keys_sorted = p->sort_depth ? B.keysB : nullptr;   // the grouping alone has no keys
    return B2S_OK;
  }
  unsigned long long* kA = B.keysA;
  unsigned long long* kB = B.keysB;
  int* vA = (passes % 2 == 0) ? B.vals : B.valsB;
  int* vB = (passes % 2 == 0) ? B.valsB : B.vals;
  {
    StageTimer t(ctx, ST_BIN, st);
    rc = launch_bin(vp, n, max_pairs, B.rect, B.tmask, B.dbits, B.cnt, B.bsum, kA, vA, B.counters, mirror, st);
  }
  if (rc != B2S_OK) return rc;
  ticket_mark(ctx, B.counters, st);
  if (keys_unsorted_copy != nullptr)
    B2S_CUDA_TRY(cudaMemcpyAsync(keys_unsorted_copy, kA, (size_t)max_pairs * 8, cudaMemcpyDeviceToDevice, st));
  if (vals_unsorted_copy != nullptr)
    B2S_CUDA_TRY(cudaMemcpyAsync(vals_unsorted_copy, vA, (size_t)max_pairs * 4, cudaMemcpyDeviceToDevice, st));
  int in_b = 0;
  {
    StageTimer t(ctx, ST_SORT, st);
    rc = launch_sort(kA, vA, kB, vB, max_pairs, &B.counters->kept, begin_bit, end_bit, B.hist, B.hsum, &in_b, st);
  }
  if (rc != B2S_OK) return rc;
  unsigned long long* ks = in_b ? kB : kA;
  {
    StageTimer t(ctx, ST_RANGES, st);
    rc = launch_ranges(ks, &B.counters->kept, max_pairs, vp.n_tiles, B.ranges, st);
    if (rc == B2S_OK) rc = launch_units(B.ranges, vp.n_tiles, vp.seg, B.unit_cap, B.unit_start, B.units, st);
    if (rc == B2S_OK) rc = launch_udesc(B.ranges, B.unit_start, vp.n_tiles, vp.seg, B.unit_cap, B.udesc, B.counters, st);
  }
  if (rc != B2S_OK) return rc;
  if (keys_sorted != nullptr) *keys_sorted = ks;
  return B2S_OK;
}

static int check_sizes(int n, int w, int h, int64_t mp, size_t state_bytes, bool need_state, size_t ws_bytes) {
  if (n < 0 || mp < 0 || mp > 0x7fffffffLL) { set_error("bad n=%d or max_pairs=%lld", n, (long long)mp); return B2S_ERR_INVALID; }
  if (need_state && state_bytes < state_layout(n, w, h, mp).total) {
    set_error("state buffer too small: %zu < %zu", state_bytes, state_layout(n, w, h, mp).total);
    return B2S_ERR_WORKSPACE;
  }
  if (ws_bytes < work_layout(n, w, h, mp).total) {
    set_error("workspace too small: %zu < %zu", ws_bytes, work_layout(n, w, h, mp).total);
    return B2S_ERR_WORKSPACE;
  }
  return B2S_OK;
}

}  // namespace b2s

using namespace b2s;

extern "C" {

const char* b2s_last_error(void) { return g_err; }
const char* b2s_version(void) { return "b2splat 0.1 (sm_100a)"; }

b2s_ctx* b2s_create(int device) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
    set_error("b2s_create: CUDA device %d not available (%d devices)", device, count);
    return nullptr;
  }
  b2s_ctx* c = new b2s_ctx();
  c->device = device;
  return c;
}

void b2s_destroy(b2s_ctx* ctx) {
  if (ctx == nullptr) return;
  if (ctx->host_dev != nullptr) cudaFree(ctx->host_dev);
  if (ctx->probe != nullptr) cudaFreeHost(ctx->probe);
  for (auto& e : ctx->probe_ev) if (e != nullptr) cudaEventDestroy(e);
  for (auto& sp : ctx->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
  for (auto& e : ctx->pool) cudaEventDestroy(e);
  delete ctx;
}

size_t b2s_state_bytes(int n, int width, int height, int64_t max_pairs) {
  return state_layout(n, width, height, max_pairs).total;
}
size_t b2s_workspace_bytes(int n, int width, int height, int64_t max_pairs) {
  return work_layout(n, width, height, max_pairs).total;
}

int b2s_count_pairs(b2s_ctx* ctx, const b2s_params* p, const float* means, const float* scales,
                    const float* opacities, int n, int64_t* total_host, void* workspace, size_t ws_bytes,
                    void* stream) {
  if (ctx == nullptr || total_host == nullptr) { set_error("NULL argument"); return B2S_ERR_INVALID; }
  ViewParams vp;
  int rc = make_view(p, &vp);
  if (rc != B2S_OK) return rc;
  *total_host = 0;
  if (n == 0) return B2S_OK;
  rc = check_sizes(n, p->width, p->height, 0, 0, false, ws_bytes);
  if (rc != B2S_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  Bufs B = resolve(nullptr, workspace, n, p->width, p->height, 0);
  // the counters live in the (otherwise unused) head of the histogram scratch
  Counters* counters = (Counters*)B.hist;
  rc = launch_preprocess(vp, means, scales, nullptr, opacities, n, nullptr, nullptr, B.rect, B.tmask, B.dbits, B.cnt, B.bsum, nullptr, nullptr, st);
  if (rc != B2S_OK) return rc;
  rc = launch_bin(vp, n, 0x7fffffffLL, B.rect, B.tmask, B.dbits, B.cnt, B.bsum, nullptr, nullptr, counters, nullptr, st);
  if (rc != B2S_OK) return rc;
  Counters h;
  B2S_CUDA_TRY(cudaMemcpyAsync(&h, counters, sizeof(h), cudaMemcpyDeviceToHost, st));
  B2S_CUDA_TRY(cudaStreamSynchronize(st));
  *total_host = h.needed;
  return B2S_OK;
}

int b2s_forward(b2s_ctx* ctx, const b2s_params* p, const float* means, const float* scales, const float* colors,
                const float* opacities, int n, int64_t max_pairs, float* out_rgb, float* out_alpha,
                float* out_depth, void* state, size_t state_bytes, void* workspace, size_t ws_bytes, void* stream) {
  if (ctx == nullptr || state == nullptr || workspace == nullptr) { set_error("NULL argument"); return B2S_ERR_INVALID; }
  ViewParams vp;
  int rc = make_view(p, &vp);
  if (rc != B2S_OK) return rc;
  if (out_rgb == nullptr && vp.mode == B2S_MODE_SORTED) { set_error("out_rgb is NULL (only the weighted-sum mode keeps its result in the state)"); return B2S_ERR_INVALID; }
  rc = check_sizes(n, p->width, p->height, max_pairs, state_bytes, true, ws_bytes);
  if (rc != B2S_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  Bufs B = resolve(state, workspace, n, p->width, p->height, max_pairs);
  rc = run_binning(ctx, vp, p, means, scales, colors, opacities, n, max_pairs, B, nullptr, nullptr, nullptr, nullptr, nullptr, st);
  if (rc != B2S_OK) return rc;
  StageTimer t(ctx, ST_BLEND_FWD, st);
  if (vp.mode == B2S_MODE_SORTED)
    return launch_blend_sorted_fwd(vp, B.rec, B.vals, B.ranges, B.unit_start, B.partial, B.cs_total, B.rest_list, out_rgb, out_alpha, nullptr, B.counters, st);
  return launch_blend_wsum_fwd(vp, B.rec, B.vals, B.ranges, B.unit_start, B.units, B.udesc, B.counters, B.unit_cap, B.partial, out_rgb,
                               out_alpha, out_depth, B.acc, nullptr, st);
}

size_t b2s_prepared_view_bytes(int n) { return prepared_layout(n).total; }

int b2s_preprocess_views(b2s_ctx* ctx, const void* views_dev, int num_views, int sh_coeffs, const float* means,
                         const float* scales, const float* colors, const float* opacities, int n, void* prepared,
                         void* stream) {
  if (ctx == nullptr || views_dev == nullptr || means == nullptr || scales == nullptr || colors == nullptr ||
      opacities == nullptr || prepared == nullptr) { set_error("NULL argument"); return B2S_ERR_INVALID; }
  if (n < 0 || num_views < 0) { set_error("bad n or num_views"); return B2S_ERR_INVALID; }
  StageTimer t(ctx, ST_PREPROCESS, (cudaStream_t)stream);
  return launch_preprocess_views((const ViewParams*)views_dev, num_views, sh_coeffs > 0 ? sh_coeffs : 1, means, scales,
                                 colors, opacities, n, (char*)prepared, (cudaStream_t)stream);
}

int b2s_forward_prepared(b2s_ctx* ctx, const b2s_params* p, const void* prepared_view, int n, int64_t max_pairs,
                         float* out_rgb, float* out_alpha, float* out_depth, void* state, size_t state_bytes,
                         void* workspace, size_t ws_bytes, void* stream) {
  if (ctx == nullptr || prepared_view == nullptr || state == nullptr || workspace == nullptr) { set_error("NULL argument"); return B2S_ERR_INVALID; }
  ViewParams vp;
  int rc = make_view(p, &vp);
  if (rc != B2S_OK) return rc;
  if (vp.mode != B2S_MODE_WSUM) { set_error("prepared views are a weighted-sum (fit loop) path"); return B2S_ERR_UNSUPPORTED; }
  rc = check_sizes(n, p->width, p->height, max_pairs, state_bytes, true, ws_bytes);
  if (rc != B2S_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  Bufs B = resolve(state, workspace, n, p->width, p->height, max_pairs);
  use_prepared(B, prepared_view, n);
  rc = run_binning(ctx, vp, p, nullptr, nullptr, nullptr, nullptr, n, max_pairs, B, nullptr, nullptr, nullptr, nullptr, nullptr, st);
  if (rc != B2S_OK) return rc;
  StageTimer t(ctx, ST_BLEND_FWD, st);
  return launch_blend_wsum_fwd(vp, B.rec, B.vals, B.ranges, B.unit_start, B.units, B.udesc, B.counters, B.unit_cap, B.partial, out_rgb,
                               out_alpha, out_depth, B.acc, nullptr, st);
}

int b2s_backward(b2s_ctx* ctx, const b2s_params* p, const float* means, const float* scales, const float* colors,
                 const float* opacities, int n, int64_t max_pairs, const float* g_rgb, const float* g_alpha,
                 const float* g_depth, const void* state, void* workspace, size_t ws_bytes, float* grad_means,
                 float* grad_scales, float* grad_colors, float* grad_opacities, int accumulate, void* stream) {
  if (ctx == nullptr || g_rgb == nullptr || state == nullptr || workspace == nullptr || grad_means == nullptr ||
      grad_scales == nullptr || grad_opacities == nullptr) { set_error("NULL argument"); return B2S_ERR_INVALID; }
  ViewParams vp;
  int rc = make_view(p, &vp);
  if (rc != B2S_OK) return rc;
  if (vp.mode != B2S_MODE_WSUM || vp.exact_bbox || vp.style != B2S_STYLE_TORCH) {
    set_error("backward is implemented for the weighted-sum torch-style mode only");
    return B2S_ERR_UNSUPPORTED;
  }
  rc = check_sizes(n, p->width, p->height, max_pairs, 0, false, ws_bytes);
  if (rc != B2S_OK) return rc;
  if (n == 0) return B2S_OK;
  cudaStream_t st = (cudaStream_t)stream;
  Bufs B = resolve(const_cast<void*>(state), workspace, n, p->width, p->height, max_pairs);
  {
    StageTimer t(ctx, ST_BLEND_BWD, st);
    rc = launch_gacc_init(B.cmask, B.gacc, n, st);
    if (rc != B2S_OK) return rc;
    rc = launch_blend_wsum_bwd(vp, B.rec, B.vals, B.ranges, B.unit_start, B.units, B.udesc, B.counters, B.unit_cap, B.acc, g_rgb, g_alpha,
                               g_depth, nullptr, B.gbuf, B.gacc, st);
  }
  if (rc != B2S_OK) return rc;
  StageTimer t(ctx, ST_PREPROCESS_BWD, st);
  return launch_preprocess_bwd(&vp, nullptr, 1, vp.sh, means, scales, colors, opacities, n, 0, n, B.gacc, grad_means,
                               grad_scales, grad_colors, grad_opacities, accumulate, st);
}

/* ---- extension modes: rotations + EWA covariance, differentiable "over" compositing (splat2d.cu) ---------------- */
static int ext_view(const b2s_params* p, int blend, b2s_params* q, ViewParams* vp) {
  if (p == nullptr) { set_error("params is NULL"); return B2S_ERR_INVALID; }
  if (blend != B2S_BLEND_WSUM && blend != B2S_BLEND_OVER) { set_error("blend must be B2S_BLEND_WSUM or B2S_BLEND_OVER"); return B2S_ERR_INVALID; }
  *q = *p;
  q->enable_depth_sort = B2S_MODE_WSUM;       // the shared front end in its torch-style configuration; the blend is chosen by `blend`
  q->style = B2S_STYLE_TORCH;
  q->exact_bbox = 0;
  q->sort_depth = (blend == B2S_BLEND_OVER) ? 1 : 0;
  return make_view(q, vp);
}

int b2s_forward_ext(b2s_ctx* ctx, const b2s_params* p, const float* means, const float* scales, const float* rotations,
                    const float* colors, const float* opacities, int n, int64_t max_pairs, int blend, float ewa_dilation,
                    float* out_rgb, float* out_alpha, float* out_depth, void* state, size_t state_bytes,
                    void* workspace, size_t ws_bytes, void* stream) {
  if (ctx == nullptr || state == nullptr || workspace == nullptr || means == nullptr || scales == nullptr ||
      colors == nullptr || opacities == nullptr) { set_error("NULL argument"); return B2S_ERR_INVALID; }
  b2s_params q;
  ViewParams vp;
  int rc = ext_view(p, blend, &q, &vp);
  if (rc != B2S_OK) return rc;
  rc = check_sizes(n, p->width, p->height, max_pairs, state_bytes, true, ws_bytes);
  if (rc != B2S_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  Bufs B = resolve(state, workspace, n, p->width, p->height, max_pairs);
  {
    StageTimer t(ctx, ST_PREPROCESS, st);
    rc = launch_ext_preprocess(vp, means, scales, rotations, colors, opacities, n, ewa_dilation, B.rec, B.cmask, B.rect, B.tmask,
                               B.dbits, B.cnt, B.bsum, st);
  }
  if (rc != B2S_OK) return rc;
  rc = run_binning(ctx, vp, &q, means, scales, colors, opacities, n, max_pairs, B, nullptr, nullptr, nullptr, nullptr, nullptr, st, true);
  if (rc != B2S_OK) return rc;
  StageTimer t(ctx, ST_BLEND_FWD, st);
  return launch_blend_ext_fwd(vp, blend == B2S_BLEND_OVER, B.rec, B.vals, B.ranges, out_rgb, out_alpha, out_depth, B.acc, st);
}

int b2s_backward_ext(b2s_ctx* ctx, const b2s_params* p, const float* means, const float* scales, const float* rotations,
                     const float* colors, const float* opacities, int n, int64_t max_pairs, int blend, float ewa_dilation,
                     const float* g_rgb, const float* g_alpha, const float* g_depth, const void* state, void* workspace,
                     size_t ws_bytes, float* grad_means, float* grad_scales, float* grad_rotations, float* grad_colors,
                     float* grad_opacities, void* stream) {
  if (ctx == nullptr || state == nullptr || workspace == nullptr || grad_means == nullptr || grad_scales == nullptr ||
      grad_colors == nullptr || grad_opacities == nullptr || (rotations != nullptr && grad_rotations == nullptr)) {
    set_error("NULL argument");
    return B2S_ERR_INVALID;
  }
  b2s_params q;
  ViewParams vp;
  int rc = ext_view(p, blend, &q, &vp);
  if (rc != B2S_OK) return rc;
  rc = check_sizes(n, p->width, p->height, max_pairs, 0, false, ws_bytes);
  if (rc != B2S_OK) return rc;
  if (n == 0) return B2S_OK;
  cudaStream_t st = (cudaStream_t)stream;
  Bufs B = resolve(const_cast<void*>(state), workspace, n, p->width, p->height, max_pairs);
  {
    StageTimer t(ctx, ST_BLEND_BWD, st);
    rc = launch_blend_ext_bwd(vp, blend == B2S_BLEND_OVER, B.rec, B.vals, B.ranges, B.acc, g_rgb, g_alpha, g_depth, B.gacc, n, st);
  }
  if (rc != B2S_OK) return rc;
  StageTimer t(ctx, ST_PREPROCESS_BWD, st);
  return launch_ext_bwd(vp, means, scales, rotations, colors, opacities, n, ewa_dilation, B.gacc, B.cmask, grad_means, grad_scales,
                        grad_rotations, grad_colors, grad_opacities, st);
}

size_t b2s_view_block_bytes(void) { return sizeof(ViewParams); }

int b2s_pack_views(const b2s_params* params, int num_views, void* out_host) {
  if (params == nullptr || out_host == nullptr || num_views < 0) { set_error("NULL argument"); return B2S_ERR_INVALID; }
  ViewParams* out = (ViewParams*)out_host;
  for (int v = 0; v < num_views; ++v) {
    const int rc = make_view(&params[v], &out[v]);
    if (rc != B2S_OK) return rc;
  }
  return B2S_OK;
}

int b2s_backward_blend(b2s_ctx* ctx, const b2s_params* p, int n, int64_t max_pairs, const float* g_rgb,
                       const float* g_alpha, const float* g_depth, const void* state, void* workspace, size_t ws_bytes,
                       float* gacc_out, void* stream) {
  if (ctx == nullptr || g_rgb == nullptr || state == nullptr || workspace == nullptr || gacc_out == nullptr) { set_error("NULL argument"); return B2S_ERR_INVALID; }
  ViewParams vp;
  int rc = make_view(p, &vp);
  if (rc != B2S_OK) return rc;
  if (vp.mode != B2S_MODE_WSUM || vp.exact_bbox || vp.style != B2S_STYLE_TORCH) {
    set_error("backward is implemented for the weighted-sum torch-style mode only");
    return B2S_ERR_UNSUPPORTED;
  }
  rc = check_sizes(n, p->width, p->height, max_pairs, 0, false, ws_bytes);
  if (rc != B2S_OK) return rc;
  if (n == 0) return B2S_OK;
  cudaStream_t st = (cudaStream_t)stream;
  Bufs B = resolve(const_cast<void*>(state), workspace, n, p->width, p->height, max_pairs);
  StageTimer t(ctx, ST_BLEND_BWD, st);
  rc = launch_gacc_init(B.cmask, gacc_out, n, st);
  if (rc != B2S_OK) return rc;
  return launch_blend_wsum_bwd(vp, B.rec, B.vals, B.ranges, B.unit_start, B.units, B.udesc, B.counters, B.unit_cap, B.acc, g_rgb, g_alpha,
                               g_depth, nullptr, B.gbuf, gacc_out, st);
}

static int fit_backward_blend_impl(b2s_ctx* ctx, const b2s_params* p, int n, int64_t max_pairs, const float* tgt,
                           const float* mask, const uint8_t* tgt8, const uint8_t* mask8, const float* depth_gt, float w_sil, float w_depth, float scale,
                           float* loss_accum, const void* state, const void* prepared_view, void* workspace,
                           size_t ws_bytes, float* gacc_out, void* stream) {
  if (ctx == nullptr || (tgt == nullptr && tgt8 == nullptr) || loss_accum == nullptr || state == nullptr || workspace == nullptr ||
      gacc_out == nullptr) { set_error("NULL argument"); return B2S_ERR_INVALID; }
  ViewParams vp;
  int rc = make_view(p, &vp);
  if (rc != B2S_OK) return rc;
  if (vp.mode != B2S_MODE_WSUM || vp.exact_bbox || vp.style != B2S_STYLE_TORCH) {
    set_error("backward is implemented for the weighted-sum torch-style mode only");
    return B2S_ERR_UNSUPPORTED;
  }
  rc = check_sizes(n, p->width, p->height, max_pairs, 0, false, ws_bytes);
  if (rc != B2S_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  Bufs B = resolve(const_cast<void*>(state), workspace, n, p->width, p->height, max_pairs);
  if (prepared_view != nullptr) use_prepared(B, prepared_view, n);
  if (depth_gt != nullptr && w_depth != 0.0f && !p->keep_depth) {
    set_error("the depth term needs the depth plane: run the forward with params.keep_depth = 1");
    return B2S_ERR_INVALID;
  }
  const FitLossArgs fl = {tgt, mask, w_sil, scale, loss_accum, depth_gt, w_depth, tgt8, mask8};
  StageTimer t(ctx, ST_BLEND_BWD, st);
  if (n > 0) {
    rc = launch_gacc_init(B.cmask, gacc_out, n, st);
    if (rc != B2S_OK) return rc;
  }
  return launch_blend_wsum_bwd(vp, B.rec, B.vals, B.ranges, B.unit_start, B.units, B.udesc, B.counters, B.unit_cap, B.acc, nullptr, nullptr,
                               nullptr, &fl, B.gbuf, gacc_out, st);
}

int b2s_fit_backward_blend(b2s_ctx* ctx, const b2s_params* p, int n, int64_t max_pairs, const float* tgt,
                           const float* mask, const float* depth_gt, float w_sil, float w_depth, float scale,
                           float* loss_accum, const void* state, const void* prepared_view, void* workspace,
                           size_t ws_bytes, float* gacc_out, void* stream) {
  return fit_backward_blend_impl(ctx, p, n, max_pairs, tgt, mask, nullptr, nullptr, depth_gt, w_sil, w_depth, scale, loss_accum,
                                 state, prepared_view, workspace, ws_bytes, gacc_out, stream);
}

int b2s_fit_backward_blend_u8(b2s_ctx* ctx, const b2s_params* p, int n, int64_t max_pairs, const uint8_t* tgt_u8,
                              const uint8_t* mask_u8, const float* depth_gt, float w_sil, float w_depth, float scale,
                              float* loss_accum, const void* state, const void* prepared_view, void* workspace,
                              size_t ws_bytes, float* gacc_out, void* stream) {
  return fit_backward_blend_impl(ctx, p, n, max_pairs, nullptr, nullptr, tgt_u8, mask_u8, depth_gt, w_sil, w_depth, scale,
                                 loss_accum, state, prepared_view, workspace, ws_bytes, gacc_out, stream);
}

int b2s_backward_params_range(b2s_ctx* ctx, const void* views_dev, int num_views, int sh_coeffs, const float* means,
                              const float* scales, const float* colors, const float* opacities, int n, int first, int count,
                              const float* gacc_all, float* grad_means, float* grad_scales, float* grad_colors,
                              float* grad_opacities, int accumulate, void* stream) {
  if (ctx == nullptr || views_dev == nullptr || gacc_all == nullptr || grad_means == nullptr || grad_scales == nullptr ||
      grad_opacities == nullptr) { set_error("NULL argument"); return B2S_ERR_INVALID; }
  if (n < 0 || num_views < 0 || first < 0 || count < 0 || (long long)first + count > n) { set_error("bad n, num_views or range"); return B2S_ERR_INVALID; }
  StageTimer t(ctx, ST_PREPROCESS_BWD, (cudaStream_t)stream);
  return launch_preprocess_bwd(nullptr, (const ViewParams*)views_dev, num_views, sh_coeffs > 0 ? sh_coeffs : 1, means,
                               scales, colors, opacities, n, first, count, gacc_all, grad_means, grad_scales, grad_colors,
                               grad_opacities, accumulate, (cudaStream_t)stream);
}

int b2s_backward_params(b2s_ctx* ctx, const void* views_dev, int num_views, int sh_coeffs, const float* means,
                        const float* scales, const float* colors, const float* opacities, int n, const float* gacc_all,
                        float* grad_means, float* grad_scales, float* grad_colors, float* grad_opacities,
                        int accumulate, void* stream) {
  return b2s_backward_params_range(ctx, views_dev, num_views, sh_coeffs, means, scales, colors, opacities, n, 0, n > 0 ? n : 0,
                                   gacc_all, grad_means, grad_scales, grad_colors, grad_opacities, accumulate, stream);
}

int b2s_state_info(b2s_ctx* ctx, const void* state, int n, int width, int height, int64_t max_pairs,
                   int64_t* info_host, void* stream) {
  if (ctx == nullptr || state == nullptr || info_host == nullptr) { set_error("NULL argument"); return B2S_ERR_INVALID; }
  const StateLayout S = state_layout(n, width, height, max_pairs);
  Counters h;
  cudaStream_t st = (cudaStream_t)stream;
  B2S_CUDA_TRY(cudaMemcpyAsync(&h, (const char*)state + S.counters, sizeof(h), cudaMemcpyDeviceToHost, st));
  B2S_CUDA_TRY(cudaStreamSynchronize(st));
  info_host[0] = h.needed;
  info_host[1] = h.kept;
  info_host[2] = h.overflow;
  return B2S_OK;
}

// The RGBA8 paths keep their "state" inside the workspace: workspace = [work | state].
static int render_rgba8_impl(b2s_ctx* ctx, const ViewParams& vp, const b2s_params* p, const float* means, const float* scales,
                             const float* colors, const float* opac, int n, int64_t max_pairs, uint8_t* out_rgba,
                             void* workspace, size_t ws_bytes, cudaStream_t st) {
  const size_t wbytes = work_layout(n, p->width, p->height, max_pairs).total;
  const size_t sbytes = state_layout(n, p->width, p->height, max_pairs).total;
  if (ws_bytes < wbytes + sbytes) {
    set_error("rgba8 workspace too small: %zu < %zu", ws_bytes, wbytes + sbytes);
    return B2S_ERR_WORKSPACE;
  }
  Bufs B = resolve((char*)workspace + wbytes, workspace, n, p->width, p->height, max_pairs);
  int rc = run_binning(ctx, vp, p, means, scales, colors, opac, n, max_pairs, B, nullptr, nullptr, nullptr, nullptr, nullptr, st);
  if (rc != B2S_OK) return rc;
  StageTimer t(ctx, ST_BLEND_FWD, st);
  if (vp.mode == B2S_MODE_SORTED)
    return launch_blend_sorted_fwd(vp, B.rec, B.vals, B.ranges, B.unit_start, B.partial, B.cs_total, B.rest_list, nullptr, nullptr, out_rgba, B.counters, st);
  return launch_blend_wsum_fwd(vp, B.rec, B.vals, B.ranges, B.unit_start, B.units, B.udesc, B.counters, B.unit_cap, B.partial, nullptr,
                               nullptr, nullptr, nullptr, out_rgba, st);
}

int b2s_render_rgba8(b2s_ctx* ctx, const b2s_params* p, const float* means, const float* scales,
                     const float* colors, const float* opacities, int n, int64_t max_pairs, uint8_t* out_rgba,
                     void* workspace, size_t ws_bytes, void* stream) {
  if (ctx == nullptr || out_rgba == nullptr || workspace == nullptr) { set_error("NULL argument"); return B2S_ERR_INVALID; }
  ViewParams vp;
  int rc = make_view(p, &vp);
  if (rc != B2S_OK) return rc;
  if (n < 0 || max_pairs < 0 || max_pairs > 0x7fffffffLL) { set_error("bad n or max_pairs"); return B2S_ERR_INVALID; }
  return render_rgba8_impl(ctx, vp, p, means, scales, colors, opacities, n, max_pairs, out_rgba, workspace, ws_bytes,
                           (cudaStream_t)stream);
}

// Host-pointer entry: same argument meaning as gr::render_gaussians (include/gr/renderer.h:33-39).
// Like the reference's CUDA wrapper (src/renderer.cu:361-405) it uploads the four arrays, renders
// and downloads the image synchronously; unlike it, it sizes the pair buffers exactly (one count
// pass) and keeps a per-ctx grow-only device arena instead of a function-local static.
int b2s_render_rgba8_host(b2s_ctx* ctx, const b2s_params* p, const float* means_host, const float* scales_host,
                          const float* colors_host, const float* opacities_host, int n, uint8_t* out_rgba_host) {
  if (ctx == nullptr || out_rgba_host == nullptr) { set_error("NULL argument"); return B2S_ERR_INVALID; }
  ViewParams vp;
  int rc = make_view(p, &vp);
  if (rc != B2S_OK) return rc;
  if (vp.sh != 1) { set_error("the native entry takes (N,3) colours only (src/bindings.cpp:50)"); return B2S_ERR_INVALID; }
  const size_t pixels = (size_t)p->width * p->height;
  if (n <= 0) {   // src/renderer.cu:279-281
    memset(out_rgba_host, 0, pixels * 4);
    return B2S_OK;
  }
  B2S_CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t st = 0;
  const size_t in_bytes = align_up((size_t)n * 12) * 3 + align_up((size_t)n * 4);
  const size_t out_bytes = align_up(pixels * 4);
  auto ensure = [&](size_t bytes) -> int {
    if (bytes <= ctx->host_dev_bytes) return B2S_OK;
    if (ctx->host_dev != nullptr) B2S_CUDA_TRY(cudaFree(ctx->host_dev));
    ctx->host_dev = nullptr;
    ctx->host_dev_bytes = 0;
    B2S_CUDA_TRY(cudaMalloc(&ctx->host_dev, bytes));
    ctx->host_dev_bytes = bytes;
    return B2S_OK;
  };
  // pass 1: upload + count
  const size_t count_ws = work_layout(n, p->width, p->height, 0).total;
  rc = ensure(in_bytes + out_bytes + count_ws);
  if (rc != B2S_OK) return rc;
  auto carve = [&](float** m, float** s, float** c, float** o, uint8_t** img, char** ws) {
    char* base = (char*)ctx->host_dev;
    *m = (float*)base; base += align_up((size_t)n * 12);
    *s = (float*)base; base += align_up((size_t)n * 12);
    *c = (float*)base; base += align_up((size_t)n * 12);
    *o = (float*)base; base += align_up((size_t)n * 4);
    *img = (uint8_t*)base; base += out_bytes;
    *ws = base;
  };
  float *dm, *ds, *dc, *dop;
  uint8_t* dimg;
  char* ws;
  carve(&dm, &ds, &dc, &dop, &dimg, &ws);
  B2S_CUDA_TRY(cudaMemcpyAsync(dm, means_host, (size_t)n * 12, cudaMemcpyHostToDevice, st));
  B2S_CUDA_TRY(cudaMemcpyAsync(ds, scales_host, (size_t)n * 12, cudaMemcpyHostToDevice, st));
  B2S_CUDA_TRY(cudaMemcpyAsync(dop, opacities_host, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  int64_t total = 0;
  rc = b2s_count_pairs(ctx, p, dm, ds, dop, n, &total, ws, count_ws, st);
  if (rc != B2S_OK) return rc;
  if (total > 0x7fffffffLL) { set_error("too many (Gaussian,tile) pairs: %lld", (long long)total); return B2S_ERR_OVERFLOW; }
  // pass 2: render with exact capacity
  const size_t need = work_layout(n, p->width, p->height, total).total + state_layout(n, p->width, p->height, total).total;
  if (in_bytes + out_bytes + need > ctx->host_dev_bytes) {
    rc = ensure((in_bytes + out_bytes + need) * 5 / 4);
    if (rc != B2S_OK) return rc;
    carve(&dm, &ds, &dc, &dop, &dimg, &ws);
    B2S_CUDA_TRY(cudaMemcpyAsync(dm, means_host, (size_t)n * 12, cudaMemcpyHostToDevice, st));
    B2S_CUDA_TRY(cudaMemcpyAsync(ds, scales_host, (size_t)n * 12, cudaMemcpyHostToDevice, st));
    B2S_CUDA_TRY(cudaMemcpyAsync(dop, opacities_host, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  }
  B2S_CUDA_TRY(cudaMemcpyAsync(dc, colors_host, (size_t)n * 12, cudaMemcpyHostToDevice, st));
  rc = render_rgba8_impl(ctx, vp, p, dm, ds, dc, dop, n, total, dimg, ws, need, st);
  if (rc != B2S_OK) return rc;
  B2S_CUDA_TRY(cudaMemcpyAsync(out_rgba_host, dimg, pixels * 4, cudaMemcpyDeviceToHost, st));
  B2S_CUDA_TRY(cudaStreamSynchronize(st));
  return B2S_OK;
}

int b2s_dump_bins(b2s_ctx* ctx, const b2s_params* p, const float* means, const float* scales, const float* opacities,
                  int n, int64_t max_pairs, float* px, float* py, float* sx, float* sy, float* zabs, int32_t* bbox,
                  int32_t* cnt, uint64_t* keys_unsorted, int32_t* vals_unsorted, uint64_t* keys_sorted,
                  int32_t* vals_sorted, int32_t* ranges, int64_t* total, void* workspace, size_t ws_bytes, void* stream) {
  if (ctx == nullptr || workspace == nullptr) { set_error("NULL argument"); return B2S_ERR_INVALID; }
  ViewParams vp;
  int rc = make_view(p, &vp);
  if (rc != B2S_OK) return rc;
  const size_t wbytes = work_layout(n, p->width, p->height, max_pairs).total;
  const size_t sbytes = state_layout(n, p->width, p->height, max_pairs).total;
  const size_t dbg_bytes = align_up((size_t)(n > 0 ? n : 1) * 5 * 4);
  if (ws_bytes < wbytes + sbytes + dbg_bytes) {
    set_error("dump_bins workspace too small: %zu < %zu", ws_bytes, wbytes + sbytes + dbg_bytes);
    return B2S_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  Bufs B = resolve((char*)workspace + wbytes, workspace, n, p->width, p->height, max_pairs);
  float* dbg = (float*)((char*)workspace + wbytes + sbytes);
  unsigned long long* ks = nullptr;
  rc = run_binning(ctx, vp, p, means, scales, nullptr, opacities, n, max_pairs, B, dbg, bbox, &ks,
                   (unsigned long long*)keys_unsorted, vals_unsorted, st);
  if (rc != B2S_OK) return rc;
  const size_t nb = (size_t)n * 4;
  if (n > 0) {
    if (px) B2S_CUDA_TRY(cudaMemcpyAsync(px, dbg, nb, cudaMemcpyDeviceToDevice, st));
    if (py) B2S_CUDA_TRY(cudaMemcpyAsync(py, dbg + n, nb, cudaMemcpyDeviceToDevice, st));
    if (sx) B2S_CUDA_TRY(cudaMemcpyAsync(sx, dbg + 2 * (size_t)n, nb, cudaMemcpyDeviceToDevice, st));
    if (sy) B2S_CUDA_TRY(cudaMemcpyAsync(sy, dbg + 3 * (size_t)n, nb, cudaMemcpyDeviceToDevice, st));
    if (zabs) B2S_CUDA_TRY(cudaMemcpyAsync(zabs, dbg + 4 * (size_t)n, nb, cudaMemcpyDeviceToDevice, st));
    if (cnt) B2S_CUDA_TRY(cudaMemcpyAsync(cnt, B.cnt, nb, cudaMemcpyDeviceToDevice, st));
  }
  if (max_pairs > 0) {
    if (keys_sorted && ks) B2S_CUDA_TRY(cudaMemcpyAsync(keys_sorted, ks, (size_t)max_pairs * 8, cudaMemcpyDeviceToDevice, st));
    if (keys_sorted && !ks) B2S_CUDA_TRY(cudaMemsetAsync(keys_sorted, 0, (size_t)max_pairs * 8, st));   // counting path: no keys
    if (vals_sorted) B2S_CUDA_TRY(cudaMemcpyAsync(vals_sorted, B.vals, (size_t)max_pairs * 4, cudaMemcpyDeviceToDevice, st));
  }
  if (ranges) B2S_CUDA_TRY(cudaMemcpyAsync(ranges, B.ranges, (size_t)vp.n_tiles * 8, cudaMemcpyDeviceToDevice, st));
  if (total) B2S_CUDA_TRY(cudaMemcpyAsync(total, &B.counters->needed, 8, cudaMemcpyDeviceToDevice, st));
  return B2S_OK;
}

size_t b2s_sort_tmp_bytes(int64_t m) {
  const size_t mp = (size_t)(m > 0 ? m : 1);
  const size_t nb = (mp + SORT_KPB - 1) / SORT_KPB;
  return align_up(mp * 8) * 2 + align_up(mp * 4) * 2 + align_up(nb * 256 * 4) +
         align_up(((nb * 256 + 4095) / 4096 + 1) * 4) + 256;
}

int b2s_sort_pairs(b2s_ctx* ctx, const uint64_t* keys_in, const int32_t* vals_in, uint64_t* keys_out,
                   int32_t* vals_out, int64_t m, int begin_bit, int end_bit, void* tmp, size_t tmp_bytes,
                   void* stream) {
  if (ctx == nullptr || tmp == nullptr) { set_error("NULL argument"); return B2S_ERR_INVALID; }
  if (m < 0 || m > 0x7fffffffLL || begin_bit < 0 || end_bit > 64 || begin_bit > end_bit) { set_error("bad sort arguments"); return B2S_ERR_INVALID; }
  if (tmp_bytes < b2s_sort_tmp_bytes(m)) { set_error("sort tmp too small"); return B2S_ERR_WORKSPACE; }
  if (m == 0) return B2S_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t mp = (size_t)m;
  const size_t nb = (mp + SORT_KPB - 1) / SORT_KPB;
  char* q = (char*)tmp;
  unsigned long long* kA = (unsigned long long*)q; q += align_up(mp * 8);
  unsigned long long* kB = (unsigned long long*)q; q += align_up(mp * 8);
  int* vA = (int*)q; q += align_up(mp * 4);
  int* vB = (int*)q; q += align_up(mp * 4);
  int* hist = (int*)q; q += align_up(nb * 256 * 4);
  int* hsum = (int*)q; q += align_up(((nb * 256 + 4095) / 4096 + 1) * 4);
  int* count_dev = (int*)q;
  const int mi = (int)m;
  B2S_CUDA_TRY(cudaMemcpyAsync(count_dev, &mi, 4, cudaMemcpyHostToDevice, st));
  B2S_CUDA_TRY(cudaMemcpyAsync(kA, keys_in, mp * 8, cudaMemcpyDeviceToDevice, st));
  B2S_CUDA_TRY(cudaMemcpyAsync(vA, vals_in, mp * 4, cudaMemcpyDeviceToDevice, st));
  int in_b = 0;
  int rc = launch_sort(kA, vA, kB, vB, m, count_dev, begin_bit, end_bit, hist, hsum, &in_b, st);
  if (rc != B2S_OK) return rc;
  B2S_CUDA_TRY(cudaMemcpyAsync(keys_out, in_b ? kB : kA, mp * 8, cudaMemcpyDeviceToDevice, st));
  B2S_CUDA_TRY(cudaMemcpyAsync(vals_out, in_b ? vB : vA, mp * 4, cudaMemcpyDeviceToDevice, st));
  return B2S_OK;
}

int b2s_fit_loss(b2s_ctx* ctx, const float* rgb, const float* alpha, const float* tgt, const float* mask, int width,
                 int height, float w_sil, float scale, float* g_rgb, float* g_alpha, float* loss_accum, void* stream) {
  if (ctx == nullptr || rgb == nullptr || tgt == nullptr || g_rgb == nullptr || loss_accum == nullptr) { set_error("NULL argument"); return B2S_ERR_INVALID; }
  if (mask != nullptr && (alpha == nullptr || g_alpha == nullptr)) { set_error("mask given without alpha/g_alpha"); return B2S_ERR_INVALID; }
  StageTimer t(ctx, ST_LOSS, (cudaStream_t)stream);
  return launch_fit_loss(rgb, alpha, tgt, mask, width, height, w_sil, scale, g_rgb, g_alpha, loss_accum, (cudaStream_t)stream);
}

int b2s_u8_to_f32(b2s_ctx* ctx, const uint8_t* src, float* dst, int64_t count, void* stream) {
  if (ctx == nullptr || src == nullptr || dst == nullptr) { set_error("NULL argument"); return B2S_ERR_INVALID; }
  if (count < 0) { set_error("bad count"); return B2S_ERR_INVALID; }
  return launch_u8_to_f32(src, dst, count, (cudaStream_t)stream);
}

int b2s_adam_step(b2s_ctx* ctx, float* params, const float* grads, float* m, float* v, int64_t count, int step,
                  float lr, float beta1, float beta2, float eps, int64_t scales_begin, int64_t scales_end,
                  float reg_scale, int64_t opac_begin, int64_t opac_end, float reg_opacity, void* stream) {
  if (ctx == nullptr || params == nullptr || grads == nullptr || m == nullptr || v == nullptr) { set_error("NULL argument"); return B2S_ERR_INVALID; }
  if (step < 1) { set_error("Adam step is 1-based"); return B2S_ERR_INVALID; }
  StageTimer t(ctx, ST_ADAM, (cudaStream_t)stream);
  return launch_adam(params, grads, m, v, count, step, lr, beta1, beta2, eps, scales_begin, scales_end, reg_scale,
                     opac_begin, opac_end, reg_opacity, nullptr, nullptr, (cudaStream_t)stream);
}

int b2s_adam_step_guarded(b2s_ctx* ctx, float* params, const float* grads, float* m, float* v, int64_t count, int step,
                          float lr, float beta1, float beta2, float eps, int64_t scales_begin, int64_t scales_end,
                          float reg_scale, int64_t opac_begin, int64_t opac_end, float reg_opacity,
                          const float* skip_flag, int* skipped_count, void* stream) {
  if (ctx == nullptr || params == nullptr || grads == nullptr || m == nullptr || v == nullptr) { set_error("NULL argument"); return B2S_ERR_INVALID; }
  if (step < 1) { set_error("Adam step is 1-based"); return B2S_ERR_INVALID; }
  StageTimer t(ctx, ST_ADAM, (cudaStream_t)stream);
  return launch_adam(params, grads, m, v, count, step, lr, beta1, beta2, eps, scales_begin, scales_end, reg_scale,
                     opac_begin, opac_end, reg_opacity, skip_flag, skipped_count, (cudaStream_t)stream);
}

size_t b2s_densify_workspace_bytes(int n) { return densify_workspace_bytes(n) + 256; }

int b2s_adam_step_multimem(b2s_ctx* ctx, float* params_mc, const float* grads_mc, const float* params_local, float* m,
                           float* v, int64_t count, int rank, int world, int step, float lr, float beta1, float beta2,
                           float eps, int64_t scales_begin, int64_t scales_end, float reg_scale, int64_t opac_begin,
                           int64_t opac_end, float reg_opacity, const float* skip_flag, int* skipped_count, void* stream) {
  if (ctx == nullptr || params_mc == nullptr || grads_mc == nullptr || params_local == nullptr || m == nullptr || v == nullptr ||
      rank < 0 || rank >= world) { set_error("NULL / bad argument"); return B2S_ERR_INVALID; }
  StageTimer t(ctx, ST_ADAM, (cudaStream_t)stream);
  return launch_adam_multimem(params_mc, grads_mc, params_local, m, v, count, rank, world, step, lr, beta1, beta2, eps,
                              scales_begin, scales_end, reg_scale, opac_begin, opac_end, reg_opacity, skip_flag,
                              skipped_count, (cudaStream_t)stream);
}

int b2s_multimem_share(int64_t count, int rank, int world, int64_t* lo, int64_t* hi) {
  if (count < 0 || world <= 0 || rank < 0 || rank >= world || lo == nullptr || hi == nullptr) { set_error("bad argument"); return B2S_ERR_INVALID; }
  long long a, b;
  multimem_share(count, rank, world, &a, &b);
  *lo = a;
  *hi = b;
  return B2S_OK;
}

int b2s_reduce_tail_multimem(b2s_ctx* ctx, const float* tail_mc, float* tail_out, int count, void* stream) {
  if (ctx == nullptr || tail_mc == nullptr || tail_out == nullptr) { set_error("NULL argument"); return B2S_ERR_INVALID; }
  return launch_tail_multimem(tail_mc, tail_out, count, (cudaStream_t)stream);
}

int b2s_densify_prune(b2s_ctx* ctx, const float* means, const float* scales_raw, const float* opacities_raw,
                      const float* colors, int n, int color_floats, int max_gaussians, double densify_ratio,
                      float prune_opacity, uint64_t seed, uint64_t iteration, float* out_means, float* out_scales_raw,
                      float* out_opacities_raw, float* out_colors, int* n_new_host, void* workspace, size_t ws_bytes,
                      void* stream) {
  if (ctx == nullptr || means == nullptr || scales_raw == nullptr || opacities_raw == nullptr || colors == nullptr ||
      out_means == nullptr || out_scales_raw == nullptr || out_opacities_raw == nullptr || out_colors == nullptr ||
      n_new_host == nullptr || workspace == nullptr) { set_error("NULL argument"); return B2S_ERR_INVALID; }
  if (n < 0 || color_floats <= 0 || max_gaussians < 0) { set_error("bad n / color_floats / max_gaussians"); return B2S_ERR_INVALID; }
  if (ws_bytes < b2s_densify_workspace_bytes(n)) { set_error("densify workspace too small"); return B2S_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  int* n_new_dev = (int*)((char*)workspace + densify_workspace_bytes(n));
  int rc = launch_densify_prune(means, scales_raw, opacities_raw, colors, n, color_floats, max_gaussians, densify_ratio,
                                prune_opacity, seed, iteration, out_means, out_scales_raw, out_opacities_raw, out_colors,
                                n_new_dev, workspace, st);
  if (rc != B2S_OK) return rc;
  B2S_CUDA_TRY(cudaMemcpyAsync(n_new_host, n_new_dev, 4, cudaMemcpyDeviceToHost, st));
  B2S_CUDA_TRY(cudaStreamSynchronize(st));
  return B2S_OK;
}

int64_t b2s_last_ticket(b2s_ctx* ctx) { return ctx == nullptr ? -1 : (int64_t)ctx->tickets - 1; }

int b2s_ticket_info(b2s_ctx* ctx, int64_t ticket, int64_t* info_host) {
  if (ctx == nullptr || info_host == nullptr) { set_error("NULL argument"); return B2S_ERR_INVALID; }
  if (ticket < 0 || ticket >= ctx->tickets || ticket + B2S_TICKET_RING < ctx->tickets) {
    set_error("ticket %lld is not among the last %d of this ctx (issued: %lld)", (long long)ticket, B2S_TICKET_RING, ctx->tickets);
    return B2S_ERR_INVALID;
  }
  const int slot = (int)(ticket % B2S_TICKET_RING);
  if (ctx->probe == nullptr || ctx->probe_ev[slot] == nullptr) { set_error("ticket ring unavailable (pinned allocation failed)"); return B2S_ERR_CUDA; }
  B2S_CUDA_TRY(cudaEventSynchronize(ctx->probe_ev[slot]));   // the binning of that call only, not the stream
  const Counters c = ctx->probe[slot];
  info_host[0] = c.needed;
  info_host[1] = c.kept;
  info_host[2] = c.overflow;
  return B2S_OK;
}

int64_t b2s_launch_count(void) { return (int64_t)g_launches.load(); }
void b2s_path_counts(int64_t out[4]) {
  if (out == nullptr) return;
  for (int i = 0; i < PATH_COUNT; ++i) out[i] = (int64_t)g_paths[i].load();
}
int b2s_sm_count(void) { return sm_count(); }
int b2s_num_stages(void) { return ST_COUNT; }
const char* b2s_stage_name(int stage) {
  static const char* names[ST_COUNT] = {"preprocess", "bin", "sort", "ranges", "blend_fwd", "loss",
                                        "blend_bwd", "preprocess_bwd", "adam"};
  return (stage >= 0 && stage < ST_COUNT) ? names[stage] : "?";
}
int b2s_timing_enable(b2s_ctx* ctx, int on) {
  if (ctx == nullptr) { set_error("NULL ctx"); return B2S_ERR_INVALID; }
  ctx->timing = on != 0;
  return B2S_OK;
}
int b2s_timing_read(b2s_ctx* ctx, float* ms_per_stage, int64_t* spans_per_stage) {
  if (ctx == nullptr || ms_per_stage == nullptr || spans_per_stage == nullptr) { set_error("NULL argument"); return B2S_ERR_INVALID; }
  for (int s = 0; s < ST_COUNT; ++s) { ms_per_stage[s] = 0.f; spans_per_stage[s] = 0; }
  if (!ctx->spans.empty()) B2S_CUDA_TRY(cudaEventSynchronize(ctx->spans.back().b));
  for (auto& sp : ctx->spans) {
    float ms = 0.f;
    B2S_CUDA_TRY(cudaEventSynchronize(sp.b));
    B2S_CUDA_TRY(cudaEventElapsedTime(&ms, sp.a, sp.b));
    ms_per_stage[sp.stage] += ms;
    spans_per_stage[sp.stage] += 1;
    ctx->pool.push_back(sp.a);
    ctx->pool.push_back(sp.b);
  }
  ctx->spans.clear();
  return B2S_OK;
}

}  // extern "C"
