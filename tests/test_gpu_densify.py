"""GPU tests of the device-side densify/prune compaction (csrc/densify.cu) against the golden
vectors of the unmodified reference function and the CPU oracle (oracle/fit_oracle.py).
Survivors must match bit for bit and in order; clones must be the same SET (torch.topk orders
them by value, the kernel by source index); the jitter is a different generator (Philox), so it
is checked statistically."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden
from gpu_util import dev, pkg
from oracle import fit_oracle

pytestmark = pytest.mark.gpu


def _run(means, scales_raw, op_raw, colors, max_g, ratio, thr, seed=7, iteration=3):
    capi = pkg("capi")
    d = dev()
    n = means.shape[0]
    cf = int(np.prod(colors.shape[1:]))
    cap = max(max_g, n, 1)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(d)
    m, s, o, c = t(means), t(scales_raw), t(op_raw), t(colors.reshape(n, cf))
    om = torch.zeros((cap, 3), device=d); os_ = torch.zeros((cap, 3), device=d)
    oo = torch.zeros(cap, device=d); oc = torch.zeros((cap, cf), device=d)
    L = capi.lib()
    wsb = L.b2s_densify_workspace_bytes(n)
    ws = torch.empty(wsb, dtype=torch.uint8, device=d)
    n_new = C.c_int(0)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    capi.check(L.b2s_densify_prune(capi.ctx(0), m.data_ptr(), s.data_ptr(), o.data_ptr(), c.data_ptr(), n, cf, max_g,
                                   ratio, thr, seed, iteration, om.data_ptr(), os_.data_ptr(), oo.data_ptr(),
                                   oc.data_ptr(), C.byref(n_new), ws.data_ptr(), wsb, st))
    k = n_new.value
    return om[:k].cpu().numpy(), os_[:k].cpu().numpy(), oo[:k].cpu().numpy(), oc[:k].cpu().numpy()


def _rows_sorted(s, o, c):
    rows = np.concatenate([o[:, None], s, c], axis=1)
    order = np.lexsort(rows.T[::-1])
    return rows[order], order


def _check(g_means, g_scales, g_op, g_colors, max_g, ratio, thr):
    n = g_means.shape[0]
    cf = int(np.prod(g_colors.shape[1:]))
    t = lambda a: torch.from_numpy(a)
    rm, rs, ro, rc, keep, src = fit_oracle.densify_and_prune(t(g_means), t(g_scales), t(g_op), t(g_colors), max_g, ratio, thr)
    n1, add = int(keep.sum()), int(src.shape[0])
    m, s, o, c = _run(g_means, g_scales, g_op, g_colors, max_g, ratio, thr)
    assert m.shape[0] == n1 + add
    # survivors: bit-exact, order preserved
    assert np.array_equal(m[:n1], rm[:n1].numpy())
    assert np.array_equal(s[:n1], rs[:n1].numpy())
    assert np.array_equal(o[:n1], ro[:n1].numpy())
    assert np.array_equal(c[:n1], rc[:n1].numpy().reshape(n1, cf))
    if add == 0:
        return
    # clones: same set of (op_raw - 0.1, scales_raw, colours) rows
    got_rows, got_order = _rows_sorted(s[n1:], o[n1:], c[n1:])
    ref_rows, ref_order = _rows_sorted(rs[n1:].numpy(), ro[n1:].numpy(), rc[n1:].numpy().reshape(add, cf))
    assert np.array_equal(got_rows, ref_rows)
    # jitter: (clone mean - source mean) / (0.25 * scale) ~ N(0,1)
    src_means = rm[:n1].numpy()[src.numpy()][ref_order]
    src_scale = (torch.nn.functional.softplus(rs[:n1][src]) + 1e-3).numpy()[ref_order]
    z = (m[n1:][got_order] - src_means) / (0.25 * src_scale)
    assert np.isfinite(z).all()
    if z.size >= 150:
        assert abs(z.mean()) <= 5.0 / np.sqrt(z.size)
        assert abs(z.std() - 1.0) <= 0.15
        assert np.abs(z).max() < 6.0
        assert abs(np.corrcoef(z[:, 0], z[:, 1])[0, 1]) < 0.25


@pytest.mark.parametrize("name", golden_names("densify_"))
def test_densify_matches_reference_golden(name):
    g = load_golden(name)
    _check(g["means"], g["scales_raw"], g["op_raw"], g["colors"], int(g["max_gaussians"]), float(g["ratio"]),
           float(g["prune_opacity"]))
    # the reference's own output: survivors identical
    n1 = int((1.0 / (1.0 + np.exp(-g["op_raw"].astype(np.float64))) > float(g["prune_opacity"])).sum())
    if n1 >= 64:
        m, s, o, c = _run(g["means"], g["scales_raw"], g["op_raw"], g["colors"], int(g["max_gaussians"]),
                          float(g["ratio"]), float(g["prune_opacity"]))
        assert m.shape[0] == g["out_means"].shape[0]
        assert np.array_equal(o[:n1], g["out_op_raw"][:n1])


@pytest.mark.parametrize("n,sh,max_g,ratio,thr", [(1, 1, 10, 0.5, 0.05), (63, 1, 100, 0.15, 0.5), (5000, 16, 20000, 0.15, 0.05),
                                                   (200000, 4, 210000, 0.15, 0.05), (4096, 1, 4096, 0.15, 0.0)])
def test_densify_random(n, sh, max_g, ratio, thr):
    r = np.random.RandomState(n)
    means = ((r.rand(n, 3) - 0.5) * 1.2).astype(np.float32)
    scales_raw = (-2.2 + 0.5 * r.randn(n, 3)).astype(np.float32)
    op_raw = (-1.0 + 1.5 * r.randn(n)).astype(np.float32)
    colors = (0.1 * r.rand(n, 3)).astype(np.float32) if sh == 1 else (0.1 * r.randn(n, sh, 3)).astype(np.float32)
    _check(means, scales_raw, op_raw, colors, max_g, ratio, thr)


def test_densify_ties_and_determinism():
    n = 3000
    r = np.random.RandomState(5)
    means = r.rand(n, 3).astype(np.float32)
    scales_raw = r.randn(n, 3).astype(np.float32)
    op_raw = np.round(r.randn(n) * 2).astype(np.float32) / 2      # heavy ties at the top-k threshold
    colors = r.rand(n, 3).astype(np.float32)
    a = _run(means, scales_raw, op_raw, colors, 4000, 0.15, 0.05, seed=1, iteration=9)
    b = _run(means, scales_raw, op_raw, colors, 4000, 0.15, 0.05, seed=1, iteration=9)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)                                 # same seed/iteration: bit-identical (rank consistency)
    c = _run(means, scales_raw, op_raw, colors, 4000, 0.15, 0.05, seed=1, iteration=10)
    assert not np.array_equal(a[0], c[0]) and np.array_equal(a[2], c[2])
    # count and multiset of clone opacities equal the oracle's despite ties
    t = lambda x: torch.from_numpy(x)
    rm, rs, ro, rc, keep, src = fit_oracle.densify_and_prune(t(means), t(scales_raw), t(op_raw), t(colors), 4000, 0.15, 0.05)
    assert a[2].shape[0] == ro.shape[0]
    assert np.array_equal(np.sort(a[2]), np.sort(ro.numpy()))


def test_fit_driver_densify_prune_step():
    """FitDriver.densify_prune rebuilds the flat buffers, resets Adam and keeps stepping."""
    import scenes
    fit = pkg("fit")
    n, V, W, H = 400, 2, 48, 32
    means, scales, colors, opac = scenes.make_scene(3, n, sh=4, s_lo=0.03, s_hi=0.15)
    cams = [tuple(a.reshape(-1).tolist() for a in scenes.orbit_camera(i, V, W, H)) for i in range(V)]
    d = fit.FitDriver(n, 4, W, H, cams, dev())
    t = lambda a: torch.from_numpy(a).to(dev())
    sr = np.log(np.expm1(np.maximum(scales - 1e-3, 1e-4))).astype(np.float32)
    orr = np.log(opac / (1 - opac)).astype(np.float32)
    d.set_params(t(means), t(sr), t(orr), t(colors))
    d.plan()
    rng = np.random.RandomState(0)
    d.set_targets({i: t(rng.rand(H, W, 3).astype(np.float32)) for i in range(V)}, None)
    for _ in range(3):
        d.step()
    p_before = d.opacities_raw().clone()
    n_new = d.densify_prune(iteration=3, max_gaussians=3000, densify_ratio=0.15, prune_opacity=0.05, seed=11)
    n1 = int((torch.sigmoid(p_before) > 0.05).sum())
    assert n_new == n1 + min(3000 - n1, int(n1 * 0.15)) and d.n == n_new
    assert d.step_no == 0 and not d.m.any() and not d.v.any()       # Adam state reset (fit_multiview_stub.py:319-325)
    l0 = float(d.step().item())
    for _ in range(10):
        l1 = float(d.step().item())
    assert np.isfinite(l1) and l1 < l0
    assert not d.check_overflow()
