/* CPU ORACLE (test infrastructure, NOT product code) -- projection / bbox / tile
 * binning / 64-bit key / stable sort / tile ranges.
 *
 * Plain-C restatement of the integer side of the render path, against which the
 * CUDA kernels must be BIT-EXACT.  Build: gcc -O2 -ffp-contract=off (no FMA
 * contraction; every + and * below is one IEEE-754 binary32 rounding, evaluated
 * left to right exactly as written).
 *
 * What the reference pins (by code, it has no tests):
 *   - row-major mat4 * vec4, left-to-right sums
 *       /root/reference/src/renderer_cpu.cpp:21-26, src/renderer.cu:18-25
 *   - projection, validity, sigma, pixel bbox  (style 0 = torch renderer
 *       /root/reference/python/torch_renderer.py:57-78,147-150 ; style 1 = C++
 *       renderers /root/reference/src/renderer_cpu.cpp:166-200, src/renderer.cu:54-84)
 *   - depth order = camera-space z descending
 *       /root/reference/src/renderer_cpu.cpp:138-146
 * What it does NOT contain (tiles, keys, ranges): "parity unpinned" against the
 * reference; this file is the definition the CUDA path is held to.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may load this.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static void mat4_vec4(const float m[16], float a, float b, float c, float d, float out[4]) {
  for (int r = 0; r < 4; ++r) {
    float t = m[4 * r + 0] * a;
    t = t + m[4 * r + 1] * b;
    t = t + m[4 * r + 2] * c;
    t = t + m[4 * r + 3] * d;
    out[r] = t;
  }
}

static uint32_t f32_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* ascending key order == descending camera z (closest first), ties keep index order */
uint32_t b2o_depth_bits(float zcam) {
  uint32_t u = f32_bits(zcam);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);   /* ascending-order-preserving */
  return ~u;
}

/* Kept tiles of a rect of at most 8 x 8 tiles: bit (ty-ty0)*w + (tx-tx0).  With cull != 0 a tile is kept iff its
 * pixel centre nearest to the Gaussian lies inside the k-sigma ellipse (the torch-style weighted sum only; every
 * pixel of a dropped tile has weight < op*exp(-k^2/2)).  Mirrors tile_cull_mask() of csrc/common.cuh op for op. */
uint64_t b2o_tile_mask(float px, float py, float sx, float sy, float k, int tx0, int ty0, int w, int h, int tile,
                       int cull) {
  const float isx = 1.0f / sx, isy = 1.0f / sy, kk = k * k;
  float uxx[8];
  for (int c = 0; c < 8; ++c) {
    const float lo = (float)((tx0 + c) * tile) + 0.5f, hi = lo + (float)(tile - 1);
    const float u = (fminf(fmaxf(px, lo), hi) - px) * isx;
    uxx[c] = u * u;
  }
  uint64_t m = 0;
  for (int r = 0; r < h; ++r) {
    const float lo = (float)((ty0 + r) * tile) + 0.5f, hi = lo + (float)(tile - 1);
    const float u = (fminf(fmaxf(py, lo), hi) - py) * isy;
    const float uyy = u * u;
    for (int c = 0; c < w; ++c)
      if (!cull || uxx[c] + uyy <= kk) m |= (uint64_t)1 << (r * w + c);
  }
  return m;
}

/* Per-Gaussian projection + bbox + tile rect.  Arrays are caller-allocated, length n
 * (bbox/rect: 4n, xmin ymin xmax ymax / tx0 ty0 tx1 ty1, inclusive).  cnt[i] = number
 * of tiles touched (0 when culled).  Returns the total number of (Gaussian,tile) pairs. */
int64_t b2o_project(const float* means, const float* scales, const float* opac,
                    const float* view, const float* proj, int n, int width, int height,
                    float k, int tile, int style,
                    float* px_o, float* py_o, float* sx_o, float* sy_o, float* zabs_o,
                    float* zcam_o, int32_t* bbox, int32_t* rect, int32_t* cnt, uint64_t* tmask, int cull) {
  const float fx = fabsf(proj[0]), fy = fabsf(proj[5]);
  const float wm1 = (float)(width - 1), hm1 = (float)(height - 1);
  int64_t total = 0;
  for (int i = 0; i < n; ++i) {
    float cam[4], clip[4];
    mat4_vec4(view, means[3 * i], means[3 * i + 1], means[3 * i + 2], 1.0f, cam);
    mat4_vec4(proj, cam[0], cam[1], cam[2], cam[3], clip);
    const float w = clip[3];
    float nx, ny, nz, zabs, ssx, ssy;
    int ok;
    if (style == 0) {            /* torch_renderer.py:66-76 */
      const float ws = (fabsf(w) < 1e-8f) ? 1.0f : w;
      nx = clip[0] / ws; ny = clip[1] / ws; nz = clip[2] / ws;
      ok = (nz >= -1.0f) && (nz <= 1.0f) && (w != 0.0f);
      zabs = fmaxf(fabsf(cam[2]), 1e-6f);
      ssx = fabsf(scales[3 * i]); ssy = fabsf(scales[3 * i + 1]);
      ok = ok && (opac[i] >= 0.0f);           /* clamp_min(0): op<0 has weight 0 and gradient 0; op==0 keeps dL/dop */
    } else {                     /* renderer_cpu.cpp:172-181 */
      const float iw = 1.0f / ((w == 0.0f) ? 1.0f : w);
      nx = clip[0] * iw; ny = clip[1] * iw; nz = clip[2] * iw;
      ok = (w != 0.0f) && !(nz < -1.0f || nz > 1.0f) && (nz == nz);
      zabs = fabsf(cam[2]) + 1e-6f;
      ssx = scales[3 * i]; ssy = scales[3 * i + 1];
      ok = ok && (opac[i] >= 1e-5f);          /* a < 1e-5 is skipped, :212 */
    }
    const float px = (nx * 0.5f + 0.5f) * wm1;
    const float py = (1.0f - (ny * 0.5f + 0.5f)) * hm1;
    float sx = ssx * 0.5f * (float)width * fx / zabs;
    float sy = ssy * 0.5f * (float)height * fy / zabs;
    sx = fmaxf(sx, 1.0f);
    sy = fmaxf(sy, 1.0f);
    const float rx = k * sx, ry = k * sy;
    const float lox = floorf(px - rx), hix = ceilf(px + rx);
    const float loy = floorf(py - ry), hiy = ceilf(py + ry);
    ok = ok && (hix >= 0.0f) && (lox <= wm1) && (hiy >= 0.0f) && (loy <= hm1);
    px_o[i] = px; py_o[i] = py; sx_o[i] = sx; sy_o[i] = sy; zabs_o[i] = zabs; zcam_o[i] = cam[2];
    if (!ok) {
      bbox[4 * i] = bbox[4 * i + 1] = 0; bbox[4 * i + 2] = bbox[4 * i + 3] = -1;
      rect[4 * i] = rect[4 * i + 1] = 0; rect[4 * i + 2] = rect[4 * i + 3] = -1;
      cnt[i] = 0;
      tmask[i] = 0;
      continue;
    }
    const int xmin = (int)fmaxf(lox, 0.0f), xmax = (int)fminf(hix, wm1);
    const int ymin = (int)fmaxf(loy, 0.0f), ymax = (int)fminf(hiy, hm1);
    bbox[4 * i] = xmin; bbox[4 * i + 1] = ymin; bbox[4 * i + 2] = xmax; bbox[4 * i + 3] = ymax;
    const int tx0 = xmin / tile, tx1 = xmax / tile, ty0 = ymin / tile, ty1 = ymax / tile;
    rect[4 * i] = tx0; rect[4 * i + 1] = ty0; rect[4 * i + 2] = tx1; rect[4 * i + 3] = ty1;
    const int tw = tx1 - tx0 + 1, th = ty1 - ty0 + 1;
    cnt[i] = tw * th;
    tmask[i] = 0;
    if (tw <= 8 && th <= 8) {      /* small rects carry an explicit tile mask */
      tmask[i] = b2o_tile_mask(px, py, sx, sy, k, tx0, ty0, tw, th, tile, cull);
      cnt[i] = __builtin_popcountll(tmask[i]);
    }
    total += cnt[i];
  }
  return total;
}

/* Emit (key,value) pairs: Gaussian i writes its tiles row-major at off[i]+j, where off
 * is the exclusive prefix sum of cnt.  key = tile_id << 32 | depth_bits(zcam). */
void b2o_emit(const int32_t* rect, const int32_t* cnt, const uint64_t* tmask, const float* zcam, int n, int tiles_x,
              uint64_t* keys, int32_t* vals) {
  int64_t o = 0;
  for (int i = 0; i < n; ++i) {
    if (cnt[i] == 0) continue;
    const uint64_t d = b2o_depth_bits(zcam[i]);
    const int tw = rect[4 * i + 2] - rect[4 * i] + 1, th = rect[4 * i + 3] - rect[4 * i + 1] + 1;
    const int masked = tw <= 8 && th <= 8;
    for (int ty = rect[4 * i + 1]; ty <= rect[4 * i + 3]; ++ty)
      for (int tx = rect[4 * i]; tx <= rect[4 * i + 2]; ++tx) {
        if (masked && !((tmask[i] >> ((ty - rect[4 * i + 1]) * tw + (tx - rect[4 * i]))) & 1)) continue;
        keys[o] = ((uint64_t)(uint32_t)(ty * tiles_x + tx) << 32) | d;
        vals[o] = i;
        ++o;
      }
  }
}

/* Stable LSD radix sort on key bits [begin_bit, end_bit), 8 bits per pass. */
void b2o_sort(uint64_t* keys, int32_t* vals, int64_t m, int begin_bit, int end_bit) {
  uint64_t* k2 = (uint64_t*)malloc((size_t)(m > 0 ? m : 1) * sizeof(uint64_t));
  int32_t* v2 = (int32_t*)malloc((size_t)(m > 0 ? m : 1) * sizeof(int32_t));
  uint64_t *ka = keys, *kb = k2;
  int32_t *va = vals, *vb = v2;
  for (int bit = begin_bit; bit < end_bit; bit += 8) {
    const int nb = (end_bit - bit) < 8 ? (end_bit - bit) : 8;
    const uint64_t mask = ((uint64_t)1 << nb) - 1;
    int64_t hist[257];
    memset(hist, 0, sizeof(hist));
    for (int64_t i = 0; i < m; ++i) hist[((ka[i] >> bit) & mask) + 1]++;
    for (int d = 0; d < 256; ++d) hist[d + 1] += hist[d];
    for (int64_t i = 0; i < m; ++i) {
      const int64_t p = hist[(ka[i] >> bit) & mask]++;
      kb[p] = ka[i]; vb[p] = va[i];
    }
    uint64_t* tk = ka; ka = kb; kb = tk;
    int32_t* tv = va; va = vb; vb = tv;
  }
  if (ka != keys) { memcpy(keys, ka, (size_t)m * sizeof(uint64_t)); memcpy(vals, va, (size_t)m * sizeof(int32_t)); }
  free(k2); free(v2);
}

/* ranges[2t], ranges[2t+1] = [start,end) of tile t in the sorted pair list (0,0 if empty). */
void b2o_ranges(const uint64_t* keys, int64_t m, int n_tiles, int32_t* ranges) {
  memset(ranges, 0, (size_t)n_tiles * 2 * sizeof(int32_t));
  for (int64_t i = 0; i < m; ++i) {
    const uint32_t t = (uint32_t)(keys[i] >> 32);
    if (i == 0 || (uint32_t)(keys[i - 1] >> 32) != t) ranges[2 * t] = (int32_t)i;
    if (i == m - 1 || (uint32_t)(keys[i + 1] >> 32) != t) ranges[2 * t + 1] = (int32_t)(i + 1);
  }
}

/* Scalar weighted-sum blend over the binned lists (what the CUDA WSUM kernels
 * evaluate: every Gaussian on every pixel of every tile it is binned to).  Double
 * accumulation; used as a float-tolerance cross-check and as the CPU-baseline "port"
 * of the blend for bench.py.  torch_renderer.py:181-202. */
void b2o_blend_wsum(const float* px, const float* py, const float* sx, const float* sy,
                    const float* zabs, const float* op, const float* col /* n*3, clamped */,
                    const int32_t* vals, const int32_t* ranges, int width, int height, int tile,
                    const float* bg, float* rgb, float* alpha, float* depth) {
  const int tiles_x = (width + tile - 1) / tile;
  for (int y = 0; y < height; ++y)
    for (int x = 0; x < width; ++x) {
      const int t = (y / tile) * tiles_x + x / tile;
      double A[3] = {0, 0, 0}, W = 0, D = 0;
      for (int s = ranges[2 * t]; s < ranges[2 * t + 1]; ++s) {
        const int i = vals[s];
        const double dx = (x + 0.5) - px[i], dy = (y + 0.5) - py[i];
        const double e = -0.5 * (dx * dx / ((double)sx[i] * sx[i]) + dy * dy / ((double)sy[i] * sy[i]));
        const double w = (op[i] > 0 ? op[i] : 0) * exp(e);
        W += w; D += w * zabs[i];
        A[0] += w * col[3 * i]; A[1] += w * col[3 * i + 1]; A[2] += w * col[3 * i + 2];
      }
      const int p = y * width + x;
      for (int c = 0; c < 3; ++c) {
        double v = (bg[c] + A[c]) / (1.0 + W);
        rgb[3 * p + c] = (float)(v < 0 ? 0 : (v > 1 ? 1 : v));
      }
      double a = W / (1.0 + W);
      alpha[p] = (float)(a < 0 ? 0 : (a > 1 ? 1 : a));
      double d = D / (W + 1e-6);
      depth[p] = (float)(d < 0 ? 0 : d);
    }
}
