"""Extension modes on the GPU (csrc/splat2d.cu through b2s_forward_ext / b2s_backward_ext) against
oracle/ext_oracle.py in float64: rotations + EWA covariance, differentiable front-to-back compositing.
The reference has neither: parity is against the repo's own oracle, at north_star's
tolerances -- image max-abs 1e-4, gradients relative L2 1e-3."""
import numpy as np
import pytest
import torch

import scenes
from conftest import rel_l2, report
from gpu_util import camera, dev, pkg, to_dev
from oracle import ext_oracle as ext

pytestmark = pytest.mark.gpu

IMG_TOL = 1e-4
GRAD_TOL = 1e-3
t64 = lambda a: torch.from_numpy(np.asarray(a)).to(torch.float64)


def _run(seed, n, sh, W, H, rot, blend, s_lo=0.03, s_hi=0.25, edge_cases=False, with_aux=True):
    r = pkg("renderer")
    means, scales, colors, opac = scenes.make_scene(seed, n, sh=sh, s_lo=s_lo, s_hi=s_hi, edge_cases=edge_cases)
    view, proj = scenes.orbit_camera(1, 5, W, H)
    rng = np.random.RandomState(seed + 1)
    quat = rng.randn(n, 4).astype(np.float32) if rot else None
    bg = np.array([0.1, 0.05, 0.2], np.float32)
    g_rgb = rng.randn(H, W, 3).astype(np.float32)
    g_alpha = rng.randn(H, W).astype(np.float32)
    g_depth = (0.1 * rng.randn(H, W)).astype(np.float32)
    k_ref = 3.0 if blend == "over" else None             # the over rule has an exact pixel bbox; wsum has no cutoff
    leaves_ref = [t64(a).requires_grad_(True) for a in (means, scales, colors, opac)]
    q_ref = t64(quat).requires_grad_(True) if rot else None
    rgb_ref, alpha_ref, depth_ref = ext.render_ext(leaves_ref[0], leaves_ref[1], q_ref, leaves_ref[2], leaves_ref[3],
                                                   t64(view), t64(proj), W, H, background=t64(bg), blend=blend,
                                                   cutoff_sigma=k_ref)
    loss_ref = (rgb_ref * t64(g_rgb)).sum()
    if with_aux:
        loss_ref = loss_ref + (alpha_ref * t64(g_alpha)).sum()
        if blend == "over":                               # wsum depth = D/(W+1e-6): its tails need 7 sigma
            loss_ref = loss_ref + (depth_ref * t64(g_depth)).sum()
    loss_ref.backward()

    m, s, c, o, bgd, gr, ga, gd = to_dev(means, scales, colors, opac, bg, g_rgb, g_alpha, g_depth)
    leaves = [x.requires_grad_(True) for x in (m, s, c, o)]
    q = to_dev(quat)[0].requires_grad_(True) if rot else None
    out = r.render_gaussians_torch(*leaves, camera(view, proj), W, H, background=bgd, max_gaussians=n, return_aux=with_aux,
                                   rotations=q, blend=blend)
    if with_aux:
        rgb, alpha, depth = out
        loss = (rgb * gr).sum() + (alpha * ga).sum()
        if blend == "over":
            loss = loss + (depth * gd).sum()
    else:
        rgb, alpha, depth = out, None, None
        loss = (rgb * gr).sum()
    loss.backward()
    torch.cuda.synchronize()
    res = {"rgb": float((rgb.cpu().double() - rgb_ref.detach()).abs().max())}
    if with_aux:
        res["alpha"] = float((alpha.cpu().double() - alpha_ref.detach()).abs().max())
        if blend == "over":
            res["depth_rel"] = float((depth.cpu().double() - depth_ref.detach()).abs().max() / max(float(depth_ref.max()), 1e-6))
    names = ["means", "scales", "colors", "opac"] + (["rotations"] if rot else [])
    for nm, a, b in zip(names, leaves + ([q] if rot else []), leaves_ref + ([q_ref] if rot else [])):
        res["g_" + nm] = rel_l2(a.grad.cpu().numpy(), b.grad.numpy())
    return res


def _check(res):
    assert res["rgb"] <= IMG_TOL, res
    if "alpha" in res:
        assert res["alpha"] <= IMG_TOL, res
    if "depth_rel" in res:
        assert res["depth_rel"] <= IMG_TOL, res
    for k, v in res.items():
        if k.startswith("g_"):
            assert v <= GRAD_TOL, res


@pytest.mark.parametrize("sh", [1, 4])
@pytest.mark.parametrize("rot,blend", [(False, "over"), (True, "wsum"), (True, "over")])
def test_extension_modes_match_oracle(rot, blend, sh):
    res = _run(40 + sh, 300, sh, 64, 48, rot, blend)
    report("ext_modes", rot=rot, blend=blend, sh=sh, **res)
    _check(res)


@pytest.mark.parametrize("rot,blend", [(False, "over"), (True, "wsum"), (True, "over")])
def test_extension_modes_edge_cases_and_sh16(rot, blend):
    """negative / zero opacity, negative scales, sigma clamp, off-screen and behind-the-camera Gaussians, colour clamp."""
    res = _run(77, 240, 16, 48, 40, rot, blend, edge_cases=True)
    report("ext_edge_sh16", rot=rot, blend=blend, **res)
    _check(res)


@pytest.mark.parametrize("rot,blend", [(True, "over"), (True, "wsum"), (False, "over")])
def test_extension_modes_long_lists(rot, blend):
    """tile lists of several hundred Gaussians: several shared-memory chunks per tile, early termination of
    saturated pixels in the forward and the matching start of the reverse walk in the backward."""
    res = _run(91, 2500, 1, 96, 64, rot, blend, s_lo=0.05, s_hi=0.3, with_aux=(blend == "over"))
    report("ext_long_lists", rot=rot, blend=blend, **res)
    _check(res)


def test_rgb_only_call_and_argument_checks():
    r = pkg("renderer")
    means, scales, colors, opac = scenes.make_scene(3, 50, sh=1)
    view, proj = scenes.orbit_camera(0, 3, 32, 32)
    m, s, c, o = to_dev(means, scales, colors, opac)
    img = r.render_gaussians_torch(m, s, c, o, camera(view, proj), 32, 32, blend="over")
    assert img.shape == (32, 32, 3) and bool(torch.isfinite(img).all())
    with pytest.raises(ValueError):
        r.render_gaussians_torch(m, s, c, o, camera(view, proj), 32, 32, blend="under")
    with pytest.raises(ValueError):
        r.render_gaussians_torch(m, s, c, o, camera(view, proj), 32, 32, rotations=torch.zeros(50, 3, device=dev()))
