"""GPU side of the I/O rows: 8-bit target ingestion and the headless viewer replay."""
import importlib

import numpy as np
import pytest
import torch

import scenes
from conftest import rel_l2
from gpu_util import dev, pkg
from test_gpu_fit import _driver, _setup

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("count", [1, 15, 16, 4097, 3 * 1920 * 1080])
def test_u8_to_f32_is_exact(count):
    import ctypes as C
    capi = pkg("capi")
    src = torch.randint(0, 256, (count + 3,), dtype=torch.uint8, device=dev())[3:]     # unaligned start
    for s in (src, src.clone()):
        dst = torch.empty(count, dtype=torch.float32, device=dev())
        capi.check(capi.lib().b2s_u8_to_f32(capi.ctx(dev().index), C.c_void_p(s.data_ptr()), C.c_void_p(dst.data_ptr()),
                                            count, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        # numpy's float32 true division is what the reference does (fit_multiview_stub.py:18); torch's CUDA
        # division by a scalar multiplies by the reciprocal and differs in the last bit
        assert np.array_equal(dst.cpu().numpy(), s.cpu().numpy().astype(np.float32) / np.float32(255.0))


def test_step_from_host_accepts_u8_targets():
    S = _setup(4, V=4)
    u8_t = [np.round(t * 255).astype(np.uint8) for t in S["tgts"]]
    u8_m = [np.round(m * 255).astype(np.uint8) for m in S["masks"]]
    S2 = dict(S, tgts=[t.astype(np.float32) / 255.0 for t in u8_t], masks=[m.astype(np.float32) / 255.0 for m in u8_m])
    d1, d2 = _driver(S2), _driver(S2, lanes=2)
    host_t = {i: torch.from_numpy(u8_t[i]).pin_memory() for i in range(S["V"])}
    host_m = {i: torch.from_numpy(u8_m[i]).pin_memory() for i in range(S["V"])}
    for _ in range(3):
        l1 = float(d1.step().item())
        l2 = d2.step_from_host(host_t, host_m)
        assert abs(l1 - l2) <= 1e-6
    assert rel_l2(d1.p.cpu().numpy(), d2.p.cpu().numpy()) <= 1e-4


def test_viewer_replay_matches_host_entry(tmp_path):
    io = pkg("io")
    r = pkg("renderer")
    means, scales, colors, opac = scenes.make_scene(11, 3000, sh=1, s_lo=0.01, s_hi=0.05)
    io.save_gaussians_npz(tmp_path / "g.npz", means, scales, colors, opac.reshape(-1, 1))   # (N,1) like some writers
    model = io.load_gaussians_npz(tmp_path / "g.npz")
    W, H = 160, 90
    rep = io.ViewerReplay(model, W, H, device=dev())
    frames = rep.orbit(5)
    assert len(frames) == 5 and frames[0].shape == (H, W, 4) and frames[0].dtype == np.uint8
    assert any(not np.array_equal(frames[0], f) for f in frames[1:])                      # the camera moves
    proj = io.perspective_np(60.0, W / H, 0.01, 100.0)
    for i in (0, 3):
        view = rep.view_matrix(2.0 * np.pi * i / 5)
        ref = r.render_gaussians(means, scales, colors, opac, W, H, view, proj, np.array([0.02, 0.02, 0.02], np.float32),
                                 enable_depth_sort=1)
        assert np.array_equal(frames[i], ref)          # same kernels, same inputs: the resident path is the host path
    io.save_ppm(tmp_path / "f0.ppm", frames[0])
    assert (tmp_path / "f0.ppm").stat().st_size == len(b"P6\n160 90\n255\n") + W * H * 3
