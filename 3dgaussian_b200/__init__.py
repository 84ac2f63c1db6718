"""3dgaussian_b200 -- B200-native Gaussian-splatting rasterizer behind the render entry
points of Kirkice/3DGaussian (python/torch_renderer.py, include/gr/renderer.h).

The directory name starts with a digit, so import it with
    importlib.import_module("3dgaussian_b200")
Public surface:
    renderer.render_gaussians_torch, renderer.Camera, renderer.perspective, renderer.look_at
    renderer.render_gaussians (RGBA8, numpy in/out; the pybind-compatible entry)
    fit.FitDriver (on-device multi-view fit loop)
    python/  -- drop-in `torch_renderer` / `device_utils` / `gaussian_renderer` modules that
                let the reference's fit_multiview_stub.py run unchanged
"""
__version__ = "0.1.0"
