"""Drop-in for the reference's pybind11 extension module `gaussian_renderer`
(src/bindings.cpp:27-100): one function, render_gaussians(means, scales, colors, opacities,
width=800, height=600, view, proj, background=None) -> uint8 (H,W,4)."""
from _load import renderer as _r

render_gaussians = _r.render_gaussians
