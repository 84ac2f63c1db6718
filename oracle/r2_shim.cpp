// CPU ORACLE build glue (test infrastructure, NOT product code).
//
// extern "C" entry over the UNMODIFIED reference CPU renderer so that it can be
// driven from ctypes with enable_depth_sort exposed (the reference's own pybind
// binding hides that flag: /root/reference/src/bindings.cpp:72-77).
// Compiled together with /root/reference/src/renderer_cpu.cpp where it lies
// (see oracle/Makefile); the result goes to oracle/_ref/ and is never committed.
#include <cstdint>
#include <cstring>
#include <vector>

#include "gr/renderer.h"   // /root/reference/include/gr/renderer.h:10-17

extern "C" int r2ref_render(const float* means, const float* scales, const float* colors,
                            const float* opacities, int n, int width, int height,
                            const float* view16, const float* proj16, const float* bg3,
                            int enable_depth_sort, std::uint8_t* out_rgba) {
  gr::RenderParams p;                      // include/gr/gaussian_types.h:24-46
  p.width = width;
  p.height = height;
  std::memcpy(p.view, view16, sizeof(float) * 16);
  std::memcpy(p.proj, proj16, sizeof(float) * 16);
  std::memcpy(p.background, bg3, sizeof(float) * 3);
  p.enable_depth_sort = enable_depth_sort;
  p.depth_slices = 32;
  p.force_cpu = 1;
  try {
    std::vector<std::uint8_t> img = gr::render_gaussians_cpu(means, scales, colors, opacities, n, p);
    std::memcpy(out_rgba, img.data(), img.size());
  } catch (...) {
    return -1;
  }
  return 0;
}
