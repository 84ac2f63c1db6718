"""Drop-in for the reference module of the same name (python/torch_renderer.py): put this
directory first on sys.path and `from torch_renderer import Camera, look_at, perspective,
render_gaussians_torch` (reference python/fit_multiview_stub.py:13) resolves to the
B200-native path.  Same names, signatures and defaults."""
from _load import renderer as _r

Camera = _r.Camera
perspective = _r.perspective
look_at = _r.look_at
get_default_device = _r.get_default_device
render_gaussians_torch = _r.render_gaussians_torch

__all__ = ["Camera", "perspective", "look_at", "get_default_device", "render_gaussians_torch"]
