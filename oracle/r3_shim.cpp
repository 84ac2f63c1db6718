// CPU/GPU BASELINE build glue (test + bench infrastructure, NOT product code).
//
// extern "C" entry over the UNMODIFIED reference CUDA renderer (/root/reference/src/renderer.cu:272-408,
// `gr::render_gaussians_cuda`) so that bench.py can time it on the same B200 as the same-box GPU baseline of the
// viewer config ("the GPU number to beat").  Host pointers in, RGBA8 out -- the reference's
// own calling convention, including its six blocking H2D copies and the D2H of the frame (:363-368, :405).
// Compiled with the reference source where it lies (oracle/Makefile); the result goes to oracle/_ref/.
#include <cstdint>
#include <cstring>
#include <vector>

#include "gr/renderer.h"   // /root/reference/include/gr/renderer.h:19-31

extern "C" int r3ref_render(const float* means, const float* scales, const float* colors,
                            const float* opacities, int n, int width, int height,
                            const float* view16, const float* proj16, const float* bg3,
                            int enable_depth_sort, int depth_slices, std::uint8_t* out_rgba) {
  gr::RenderParams p;                      // include/gr/gaussian_types.h:24-46
  p.width = width;
  p.height = height;
  std::memcpy(p.view, view16, sizeof(float) * 16);
  std::memcpy(p.proj, proj16, sizeof(float) * 16);
  std::memcpy(p.background, bg3, sizeof(float) * 3);
  p.enable_depth_sort = enable_depth_sort;
  p.depth_slices = depth_slices;
  p.force_cpu = 0;
  try {
    std::vector<std::uint8_t> img = gr::render_gaussians_cuda(means, scales, colors, opacities, n, p);
    std::memcpy(out_rgba, img.data(), img.size());
  } catch (...) {
    return -1;
  }
  return 0;
}
