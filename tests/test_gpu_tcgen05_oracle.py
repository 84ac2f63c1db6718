"""The shipped-default tcgen05 blend kernels against the R1 oracle where they actually run:

  * multi-step, multi-unit tiles (lists of several thousand Gaussians per tile: 128-Gaussian steps, several
    work units per tile, per-unit partial planes summed by finalize_kernel, the unit descriptor table);
  * the benchmark's own scale (BASELINE configs[3]: 1 M Gaussians, SH degree 3, 1920x1080) through a CROP oracle:
    the dense R1 restatement (oracle/r1_oracle.py, float64) evaluated on a few 32x32 windows of the frame, with
    the image cotangent zero everywhere else.

`capi.path_counts()` proves which kernel family served each call.  Tolerances are north_star's: image max-abs
1e-4, gradients relative L2 1e-3.
"""
import numpy as np
import pytest
import torch

import scenes
from conftest import rel_l2, report
from gpu_util import camera, dev, pkg, to_dev
from oracle import cpu as ocpu
from oracle import r1_oracle as r1

pytestmark = pytest.mark.gpu

IMG_TOL = 1e-4
GRAD_TOL = 1e-3


def _paths_delta(before):
    after = pkg("capi").path_counts()
    return {k: after[k] - before[k] for k in after}


@pytest.mark.parametrize("sh", [1, 4, 16])
def test_tcgen05_fwd_bwd_match_oracle_on_multi_unit_tiles(sh):
    """rgb only (no aux): tcgen05 forward AND backward, lists of > 1024 Gaussians per tile (units of 512)."""
    r, capi = pkg("renderer"), pkg("capi")
    n, W, H = 4000, 64, 48
    means, scales, colors, opac = scenes.make_scene(300 + sh, n, sh=sh, s_lo=0.05, s_hi=0.3, edge_cases=True)
    view, proj = scenes.orbit_camera(2, 5, W, H)
    bg = np.array([0.1, 0.0, 0.2], np.float32)
    bins = ocpu.bin_gaussians(means, scales, opac, view, proj, W, H, k=5.0, begin_bit=32)
    longest = int((bins["ranges"][:, 1] - bins["ranges"][:, 0]).max())
    assert longest >= 1024, longest                      # at least three 512-Gaussian units, nine 128-Gaussian steps
    rng = np.random.RandomState(3)
    g_rgb = rng.randn(H, W, 3).astype(np.float32)

    t64 = lambda a: torch.from_numpy(a).to(torch.float64)
    leaves_ref = [t64(a).requires_grad_(True) for a in (means, scales, colors, opac)]
    rgb_ref, _, _ = r1.render_r1(*leaves_ref, t64(view), t64(proj), W, H, background=t64(bg))
    (rgb_ref * t64(g_rgb)).sum().backward()

    before = capi.path_counts()
    m, s, c, o, bgd, gd = to_dev(means, scales, colors, opac, bg, g_rgb)
    leaves = [x.requires_grad_(True) for x in (m, s, c, o)]
    rgb = r.render_gaussians_torch(*leaves, camera(view, proj), W, H, background=bgd, max_gaussians=n)
    (rgb * gd).sum().backward()
    torch.cuda.synchronize()
    d = _paths_delta(before)
    assert d["fwd_tcgen05"] >= 1 and d["bwd_tcgen05"] >= 1 and d["fwd_other"] == 0 and d["bwd_other"] == 0, d

    err = float(np.abs(rgb.detach().cpu().numpy() - rgb_ref.detach().numpy()).max())
    rels = {k: rel_l2(x.grad.cpu().numpy(), y.grad.numpy())
            for k, x, y in zip(("means", "scales", "colors", "opac"), leaves, leaves_ref)}
    report("tcgen05_multi_unit", sh=sh, longest_tile_list=longest, rgb_maxabs=err, **{"grad_" + k: v for k, v in rels.items()})
    assert err <= IMG_TOL, err
    for k, v in rels.items():
        assert v <= GRAD_TOL, (k, v)


def test_tcgen05_bwd_with_alpha_gradient_matches_oracle_on_multi_unit_tiles():
    """aux outputs with a loss on rgb + alpha (no depth term): the backward is the tcgen05 kernel with g_alpha."""
    r, capi = pkg("renderer"), pkg("capi")
    n, W, H = 3000, 64, 48
    means, scales, colors, opac = scenes.make_scene(311, n, sh=4, s_lo=0.05, s_hi=0.3)
    view, proj = scenes.orbit_camera(1, 5, W, H)
    rng = np.random.RandomState(4)
    g_rgb, g_alpha = rng.randn(H, W, 3).astype(np.float32), rng.randn(H, W).astype(np.float32)
    t64 = lambda a: torch.from_numpy(a).to(torch.float64)
    leaves_ref = [t64(a).requires_grad_(True) for a in (means, scales, colors, opac)]
    rgb_ref, alpha_ref, _ = r1.render_r1(*leaves_ref, t64(view), t64(proj), W, H)
    ((rgb_ref * t64(g_rgb)).sum() + (alpha_ref * t64(g_alpha)).sum()).backward()
    before = capi.path_counts()
    m, s, c, o, gd, ga = to_dev(means, scales, colors, opac, g_rgb, g_alpha)
    leaves = [x.requires_grad_(True) for x in (m, s, c, o)]
    rgb, alpha, depth = r.render_gaussians_torch(*leaves, camera(view, proj), W, H, max_gaussians=n, return_aux=True)
    ((rgb * gd).sum() + (alpha * ga).sum()).backward()
    torch.cuda.synchronize()
    d = _paths_delta(before)
    assert d["bwd_tcgen05"] >= 1 and d["bwd_other"] == 0, d
    assert float((alpha.detach().cpu() - alpha_ref.detach()).abs().max()) <= IMG_TOL
    for k, x, y in zip(("means", "scales", "colors", "opac"), leaves, leaves_ref):
        assert rel_l2(x.grad.cpu().numpy(), y.grad.numpy()) <= GRAD_TOL, k


# ---------------------------------------------------------------- benchmark scale -----------
def _crop_oracle(means, scales, colors, opac, view, proj, W, H, windows, cot, reach=9.5):
    """Dense float64 R1 on the pixels of `windows` [(x0, y0, w, h)] only.  A Gaussian further than `reach` sigmas
    from every pixel of a window weighs < exp(-reach^2/2) ~ 2.5e-20 of its opacity there and is left out of that
    window's evaluation (its gradient from the window is that much below the kept ones).  Returns per-window
    (rgb, alpha, depth) and the gradients of sum_w <rgb_w, cot_w> wrt all four parameter arrays (zeros elsewhere)."""
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(torch.float64)
    M, S, C_, O, V, P = t(means), t(scales), t(colors), t(opac), t(view), t(proj)
    with torch.no_grad():
        px, py, zabs, valid, _ = r1.project(M, V, P, W, H)
        sx, sy = r1.screen_sigmas(S, P, zabs, W, H)
    grads = [torch.zeros_like(x) for x in (M, S, C_, O)]
    outs, sizes = [], []
    for (x0, y0, w, h), g in zip(windows, cot):
        near = (valid & (px >= x0 - reach * sx) & (px <= x0 + w + reach * sx) &
                (py >= y0 - reach * sy) & (py <= y0 + h + reach * sy))
        idx = torch.nonzero(near).reshape(-1)
        sizes.append(int(idx.numel()))
        leaves = [x[idx].clone().requires_grad_(True) for x in (M, S, C_, O)]
        iy, ix = torch.meshgrid(torch.arange(y0, y0 + h), torch.arange(x0, x0 + w), indexing="ij")
        rgb, alpha, depth = r1.render_r1(*leaves, V, P, W, H, pixels=(ix.reshape(-1), iy.reshape(-1)), chunk=1024)
        if idx.numel() > 0:                  # a background-only window has nothing to differentiate
            (rgb * t(g).reshape(-1, 3)).sum().backward()
            for acc, leaf in zip(grads, leaves):
                acc.index_add_(0, idx, leaf.grad)
        outs.append((rgb.detach().reshape(h, w, 3).numpy(), alpha.detach().reshape(h, w).numpy(),
                     depth.detach().reshape(h, w).numpy()))
    return outs, [g.numpy() for g in grads], sizes


def test_crop_oracle_at_benchmark_scale():
    """BASELINE configs[3] inputs (bench.py's seed-1234 set: 1 M Gaussians, SH degree 3) at 1920x1080, one orbit view:
    tcgen05 forward + backward through the drop-in against dense R1 on six 32x32 windows."""
    r, capi, synth = pkg("renderer"), pkg("capi"), pkg("synth")
    n, sh, W, H, V = 1_000_000, 16, 1920, 1080, 64
    means, scales, colors, opac = synth.synth_gaussians(n, sh, 1234, dev())
    view, proj = synth.orbit_camera(5, V, W, H)
    # four windows inside the projected cube (unaligned to the 16-pixel tiles), one straddling its edge, one on background
    windows = [(933, 517, 32, 32), (701, 402, 32, 32), (1180, 655, 32, 32), (1010, 300, 32, 32), (560, 540, 32, 32),
               (40, 1020, 32, 32)]
    rng = np.random.RandomState(11)
    cot = [rng.randn(h, w, 3).astype(np.float32) for (_, _, w, h) in windows]

    before = capi.path_counts()
    leaves = [x.clone().requires_grad_(True) for x in (means, scales, colors, opac)]
    rgb = r.render_gaussians_torch(*leaves, camera(view, proj), W, H, max_gaussians=n)
    loss = sum((rgb[y0:y0 + h, x0:x0 + w] * torch.from_numpy(g).to(dev())).sum() for (x0, y0, w, h), g in zip(windows, cot))
    loss.backward()
    torch.cuda.synchronize()
    d = _paths_delta(before)
    assert d["fwd_tcgen05"] >= 1 and d["bwd_tcgen05"] >= 1 and d["fwd_other"] == 0 and d["bwd_other"] == 0, d
    with torch.no_grad():
        rgb7, alpha7, depth7 = r.render_gaussians_torch(means, scales, colors, opac, camera(view, proj), W, H,
                                                        max_gaussians=n, return_aux=True)     # k = 7, depth plane

    host = [x.detach().cpu().numpy() for x in (means, scales, colors, opac)]
    outs, grads_ref, sizes = _crop_oracle(*host, view, proj, W, H, windows, cot)
    e_rgb = e_rgb7 = e_alpha = e_depth = e_depth_all = 0.0
    for (x0, y0, w, h), (o_rgb, o_alpha, o_depth) in zip(windows, outs):
        sl = (slice(y0, y0 + h), slice(x0, x0 + w))
        e_rgb = max(e_rgb, float(np.abs(rgb[sl].detach().cpu().numpy() - o_rgb).max()))
        e_rgb7 = max(e_rgb7, float(np.abs(rgb7[sl].cpu().numpy() - o_rgb).max()))
        e_alpha = max(e_alpha, float(np.abs(alpha7[sl].cpu().numpy() - o_alpha).max()))
        de = np.abs(depth7[sl].cpu().numpy() - o_depth)
        wm = o_alpha >= 1e-2 / 1.01
        e_depth = max(e_depth, float(de[wm].max()) if wm.any() else 0.0)
        e_depth_all = max(e_depth_all, float(de.max()))
    rels = {k: rel_l2(x.grad.cpu().numpy(), gr) for k, x, gr in zip(("means", "scales", "colors", "opac"), leaves, grads_ref)}
    report("crop_oracle_c4", gaussians=n, sh=sh, width=W, height=H, windows=len(windows), gaussians_per_window=sizes,
           rgb_maxabs_k5=e_rgb, rgb_maxabs_k7=e_rgb7, alpha_maxabs_k7=e_alpha, depth_maxabs_weighted_k7=e_depth,
           depth_maxabs_all_k7=e_depth_all, **{"grad_" + k: v for k, v in rels.items()})
    assert e_rgb <= IMG_TOL and e_rgb7 <= IMG_TOL and e_alpha <= IMG_TOL, (e_rgb, e_rgb7, e_alpha)
    assert e_depth <= 1e-4, (e_depth, e_depth_all)
    for k, v in rels.items():
        assert v <= GRAD_TOL, (k, v)
    # Where the truncation error actually sits at this scale (recorded, not asserted: the shipped default stays k = 5).
    # k = 5 and k = 7 agree to ~1e-8 above, i.e. the 1.5e-6 is fp16 hi/lo arithmetic, not truncation; smaller cutoffs
    # trade (k/5)^2 of the pair count against the tails:
    for kc in (4.5, 4.0):
        lv = [x.clone().requires_grad_(True) for x in (means, scales, colors, opac)]
        rk = r.render_gaussians_torch(*lv, camera(view, proj), W, H, max_gaussians=n, cutoff_sigma=kc)
        sum((rk[y0:y0 + h, x0:x0 + w] * torch.from_numpy(g).to(dev())).sum() for (x0, y0, w, h), g in zip(windows, cot)).backward()
        torch.cuda.synchronize()
        ek = max(float(np.abs(rk[y0:y0 + h, x0:x0 + w].detach().cpu().numpy() - o[0]).max()) for (x0, y0, w, h), o in zip(windows, outs))
        report("crop_oracle_c4_cutoff_sweep", cutoff_sigma=kc, rgb_maxabs=ek,
               **{"grad_" + k: rel_l2(x.grad.cpu().numpy(), gr) for k, x, gr in zip(("means", "scales", "colors", "opac"), lv, grads_ref)})


@pytest.mark.parametrize("sh", [1, 16])
def test_tcgen05_depth_plane_and_depth_gradient_match_oracle_on_multi_unit_tiles(sh):
    """return_aux=True with a loss on rgb + alpha + depth: 5-plane tcgen05 forward (depth rides on the B operand),
    depth-gradient tcgen05 backward (gD plane in a second TMEM region); k = 7."""
    r, capi = pkg("renderer"), pkg("capi")
    n, W, H = 3000, 64, 48
    means, scales, colors, opac = scenes.make_scene(320 + sh, n, sh=sh, s_lo=0.05, s_hi=0.3, edge_cases=True)
    view, proj = scenes.orbit_camera(3, 5, W, H)
    rng = np.random.RandomState(5)
    g_rgb, g_alpha = rng.randn(H, W, 3).astype(np.float32), rng.randn(H, W).astype(np.float32)
    g_depth = (0.1 * rng.randn(H, W)).astype(np.float32)
    t64 = lambda a: torch.from_numpy(a).to(torch.float64)
    leaves_ref = [t64(a).requires_grad_(True) for a in (means, scales, colors, opac)]
    rgb_ref, alpha_ref, depth_ref = r1.render_r1(*leaves_ref, t64(view), t64(proj), W, H)
    ((rgb_ref * t64(g_rgb)).sum() + (alpha_ref * t64(g_alpha)).sum() + (depth_ref * t64(g_depth)).sum()).backward()
    before = capi.path_counts()
    m, s, c, o, gd, ga, gz = to_dev(means, scales, colors, opac, g_rgb, g_alpha, g_depth)
    leaves = [x.requires_grad_(True) for x in (m, s, c, o)]
    rgb, alpha, depth = r.render_gaussians_torch(*leaves, camera(view, proj), W, H, max_gaussians=n, return_aux=True)
    ((rgb * gd).sum() + (alpha * ga).sum() + (depth * gz).sum()).backward()
    torch.cuda.synchronize()
    d = _paths_delta(before)
    assert d["fwd_tcgen05"] >= 1 and d["bwd_tcgen05"] >= 1 and d["fwd_other"] == 0 and d["bwd_other"] == 0, d
    e_rgb = float((rgb.detach().cpu() - rgb_ref.detach()).abs().max())
    e_alpha = float((alpha.detach().cpu() - alpha_ref.detach()).abs().max())
    de = (depth.detach().cpu() - depth_ref.detach()).abs().numpy()
    wm = alpha_ref.detach().numpy() >= 1e-2 / 1.01
    e_depth = float(de[wm].max())
    rels = {k: rel_l2(x.grad.cpu().numpy(), y.grad.numpy())
            for k, x, y in zip(("means", "scales", "colors", "opac"), leaves, leaves_ref)}
    report("tcgen05_depth_multi_unit", sh=sh, rgb_maxabs=e_rgb, alpha_maxabs=e_alpha, depth_maxabs_weighted=e_depth,
           depth_maxabs_all=float(de.max()), **{"grad_" + k: v for k, v in rels.items()})
    assert e_rgb <= IMG_TOL and e_alpha <= IMG_TOL and e_depth <= 1e-4, (e_rgb, e_alpha, e_depth)
    for k, v in rels.items():
        assert v <= GRAD_TOL, (k, v)
