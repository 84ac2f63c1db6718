"""CPU checks of oracle/ext_oracle.py -- the oracle of the extension modes (rotations + EWA covariance,
differentiable front-to-back compositing), which the reference does not have.  What can be
pinned is pinned here: the rotation-free modes against the R1 / R2 restatements that the reference goldens pin
(tests/test_oracle_golden.py), and the EWA projection against closed forms."""
import numpy as np
import torch

import scenes
from oracle import ext_oracle as ext
from oracle import r1_oracle as r1

t64 = lambda a: torch.from_numpy(np.asarray(a)).to(torch.float64)


def _scene(seed, n, sh, W, H, **kw):
    means, scales, colors, opac = scenes.make_scene(seed, n, sh=sh, **kw)
    view, proj = scenes.orbit_camera(1, 4, W, H)
    return [t64(a) for a in (means, scales, colors, opac)], t64(view), t64(proj)


def test_wsum_without_rotations_is_r1():
    W, H = 40, 24
    (m, s, c, o), v, p = _scene(5, 120, 4, W, H, edge_cases=True)
    bg = t64([0.1, 0.2, 0.3])
    a = ext.render_ext(m, s, None, c, o, v, p, W, H, background=bg, blend="wsum")
    b = r1.render_r1(m, s, c, o, v, p, W, H, background=bg)
    for x, y in zip(a, b):
        assert float((x - y).abs().max()) <= 1e-12


def test_over_without_rotations_is_the_sorted_mode_of_the_reference_cpu_renderer():
    """src/renderer_cpu.cpp:125-217 as restated (and pinned to goldens) by r1_oracle.render_sorted; inputs inside the
    domain where the torch-style and native-style conventions coincide (positive scales, opacity > 1e-5, colours in
    [0,1]); the ALPHA_MAX cap and |z|+1e-6 vs max(|z|,1e-6) stay below 1e-5."""
    W, H = 48, 32
    (m, s, c, o), v, p = _scene(7, 200, 1, W, H, s_lo=0.03, s_hi=0.25)
    o = o.clamp(0.05, 1.0)
    o[:5] = 1.0                                         # saturating Gaussians: a = 1 in the reference, capped here
    bg = t64([0.02, 0.02, 0.02])
    rgb, alpha, _ = ext.render_ext(m, s, None, c, o, v, p, W, H, background=bg, blend="over", cutoff_sigma=3.0)
    rgb_ref, a_ref = r1.render_sorted(m, s, c, o, v, p, W, H, background=bg, cutoff_sigma=3.0)
    assert float((rgb - rgb_ref).abs().max()) <= 1e-5
    assert float((alpha - a_ref).abs().max()) <= 1e-5


def test_ewa_on_axis_closed_form():
    """A Gaussian on the optical axis of an unrotated camera, identity quaternion: the EWA covariance is diagonal,
    cov_xx = (0.5 (W-1) P00 s_x / z)^2 + dilation -- the reference's sigma formula (torch_renderer.py:147-150) with
    W-1 for W -- and the third scale does not reach the screen."""
    W, H = 64, 48
    proj = t64(scenes.perspective(60.0, W / H, 0.01, 100.0))
    view = torch.eye(4, dtype=torch.float64)
    view[2, 3] = -3.0                                   # camera at z = +3 looking down -z
    m = t64([[0.0, 0.0, 0.0]])
    s = t64([[0.2, 0.1, 0.7]])
    q = t64([[1.0, 0.0, 0.0, 0.0]])
    px, py, z_abs, valid, _, (A, B, C), (sx, sy) = ext.splats(m, s, q, view, proj, W, H)
    assert bool(valid[0]) and abs(float(z_abs[0]) - 3.0) < 1e-12
    ex = (0.5 * (W - 1) * float(proj[0, 0]) * 0.2 / 3.0) ** 2 + ext.EWA_DILATION
    ey = (0.5 * (H - 1) * float(proj[1, 1]) * 0.1 / 3.0) ** 2 + ext.EWA_DILATION
    assert abs(float(sx[0]) ** 2 - ex) < 1e-9 and abs(float(sy[0]) ** 2 - ey) < 1e-9
    assert abs(float(B[0])) < 1e-12 and abs(float(A[0]) - 1.0 / ex) < 1e-9 and abs(float(C[0]) - 1.0 / ey) < 1e-9
    # 90 degrees about z swaps the two screen axes
    q90 = t64([[np.cos(np.pi / 4), 0.0, 0.0, np.sin(np.pi / 4)]])
    _, _, _, _, _, _, (sx2, sy2) = ext.splats(m, s, q90, view, proj, W, H)
    ex2 = (0.5 * (W - 1) * float(proj[0, 0]) * 0.1 / 3.0) ** 2 + ext.EWA_DILATION
    assert abs(float(sx2[0]) ** 2 - ex2) < 1e-9


def test_ewa_quaternion_symmetries():
    W, H = 40, 30
    (m, s, c, o), v, p = _scene(11, 60, 1, W, H)
    g = torch.Generator().manual_seed(0)
    q = torch.randn(60, 4, generator=g, dtype=torch.float64)
    a = ext.render_ext(m, s, q, c, o, v, p, W, H, blend="wsum")[0]
    b = ext.render_ext(m, s, -3.0 * q, c, o, v, p, W, H, blend="wsum")[0]     # sign and norm of q do not matter
    assert float((a - b).abs().max()) <= 1e-12
    iso = s[:, :1].expand(-1, 3).contiguous()                                  # isotropic: the rotation drops out
    a = ext.render_ext(m, iso, q, c, o, v, p, W, H, blend="over", cutoff_sigma=3.0)[0]
    b = ext.render_ext(m, iso, torch.randn(60, 4, generator=g, dtype=torch.float64), c, o, v, p, W, H, blend="over",
                       cutoff_sigma=3.0)[0]
    assert float((a - b).abs().max()) <= 1e-9


def test_over_transmittance_and_alpha_are_consistent():
    W, H = 32, 24
    (m, s, c, o), v, p = _scene(13, 80, 1, W, H)
    white = torch.ones(80, 3, dtype=torch.float64)
    rgb, alpha, depth = ext.render_ext(m, s, None, white, o, v, p, W, H, background=t64([0.0, 0.0, 0.0]), blend="over",
                                       cutoff_sigma=3.0)
    assert float((rgb[..., 0] - alpha).abs().max()) <= 1e-12                  # white Gaussians on black: rgb == alpha
    assert float(alpha.max()) <= 1.0 and float(depth.min()) >= 0.0
