"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol the header
declares, the parameter block matches, and the drop-in modules expose the reference's names.
No compute calls (there is no GPU in the build container)."""
import ctypes as C
import importlib
import inspect
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _capi():
    return importlib.import_module("3dgaussian_b200.capi")


@pytest.fixture(scope="module")
def built():
    importlib.import_module("3dgaussian_b200.build").build()
    return _capi().lib()


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "b2splat.h")).read()
    declared = set(re.findall(r"\b(b2s_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(built, name), name
    assert declared == set(_capi().EXPORTS)


def test_param_block_layout():
    capi = _capi()
    # gr::RenderParams prefix (reference include/gr/gaussian_types.h:24-46): 2 ints, 16+16+3 floats, 3 ints
    assert capi.Params.width.offset == 0 and capi.Params.height.offset == 4
    assert capi.Params.view.offset == 8 and capi.Params.proj.offset == 72
    assert capi.Params.background.offset == 136 and capi.Params.enable_depth_sort.offset == 148
    assert capi.Params.background_dev.offset == 184 and C.sizeof(capi.Params) == 200    # extension tail


def test_no_gpu_fails_loudly(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    capi = _capi()
    with pytest.raises(capi.B2SError):
        capi.ctx(0)
    r = importlib.import_module("3dgaussian_b200.renderer")
    z = torch.zeros(4, 3)
    with pytest.raises(RuntimeError):
        r.render_gaussians_torch(z, z, z, torch.zeros(4), r.Camera(torch.eye(4), torch.eye(4)), 8, 8)
    with pytest.raises(RuntimeError):
        r.get_default_device()


def test_sizes_are_monotone(built):
    a = built.b2s_workspace_bytes(1000, 128, 128, 10000)
    b = built.b2s_workspace_bytes(2000, 128, 128, 20000)
    assert 0 < a < b
    assert built.b2s_state_bytes(1000, 128, 128, 10000) < built.b2s_state_bytes(1000, 256, 256, 10000)


def test_dropin_modules_mirror_reference_names():
    shim = os.path.join(ROOT, "3dgaussian_b200", "python")
    sys.path.insert(0, shim)
    try:
        for m in ("torch_renderer", "device_utils", "gaussian_renderer"):
            sys.modules.pop(m, None)
        tr = importlib.import_module("torch_renderer")
        du = importlib.import_module("device_utils")
        gr = importlib.import_module("gaussian_renderer")
    finally:
        sys.path.remove(shim)
    for name in ("Camera", "perspective", "look_at", "render_gaussians_torch", "get_default_device"):
        assert hasattr(tr, name)
    assert hasattr(du, "get_default_device") and hasattr(gr, "render_gaussians")
    sig = inspect.signature(tr.render_gaussians_torch)
    names = list(sig.parameters)
    # reference python/torch_renderer.py:109-121
    assert names[:11] == ["means", "scales", "colors", "opacities", "camera", "width", "height", "background",
                          "max_gaussians", "chunk_size", "return_aux"]
    assert sig.parameters["max_gaussians"].default == 10000
    assert sig.parameters["chunk_size"].default == 256
    assert sig.parameters["return_aux"].default is False
    gsig = inspect.signature(gr.render_gaussians)
    assert list(gsig.parameters)[:9] == ["means", "scales", "colors", "opacities", "width", "height", "view", "proj",
                                         "background"]
    assert gsig.parameters["width"].default == 800 and gsig.parameters["height"].default == 600
    for m in ("torch_renderer", "device_utils", "gaussian_renderer", "_load"):
        sys.modules.pop(m, None)


def test_camera_helpers_match_numpy_restatement():
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import scenes
    r = importlib.import_module("3dgaussian_b200.renderer")
    p = r.perspective(60.0, 16 / 9, 0.01, 100.0).numpy()
    assert np.allclose(p, scenes.perspective(60.0, 16 / 9, 0.01, 100.0), atol=1e-6)
    eye = torch.tensor([1.0, 0.5, 2.0])
    v = r.look_at(eye, torch.zeros(3), torch.tensor([0.0, 1.0, 0.0])).numpy()
    assert np.allclose(v, scenes.look_at([1.0, 0.5, 2.0], [0, 0, 0], [0, 1, 0]), atol=1e-6)


def test_multimem_shares_tile_every_slice():
    """The partition behind b2s_adam_step_multimem (reduce-scatter + Adam + all-gather over NVLink multicast): for every
    slice length and world size the ranks' shares are disjoint, in rank order, cover the slice, and every boundary except
    the slice end is a multiple of 4 floats (16-byte multimem accesses).  Host arithmetic only: no GPU needed."""
    capi = _capi()
    for world in (1, 2, 3, 4, 8):
        for count in (0, 1, 3, 4, 5, 63, 64, 65, 1000, 4099, 3 * 250048, 48 * 250048 + 1):
            pos = 0
            for rank in range(world):
                lo, hi = capi.multimem_share(count, rank, world)
                assert lo == pos and lo <= hi <= count
                assert lo % 4 == 0 or lo == count
                pos = hi
            assert pos == count
