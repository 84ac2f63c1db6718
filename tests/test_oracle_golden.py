"""Pins the CPU oracles against the golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden, rel_l2
from oracle import cpu as ocpu
from oracle import r1_oracle as r1

R1 = golden_names("r1_")
R2 = golden_names("r2_")


def _t(a, dt):
    return torch.from_numpy(np.asarray(a)).to(dt)


def _render(g, dt, **kw):
    args = [_t(g[k], dt) for k in ("means", "scales", "colors", "opac", "view", "proj")]
    return r1.render_r1(*args, int(g["width"]), int(g["height"]), background=_t(g["bg"], dt), **kw)


def test_goldens_present():
    assert len(R1) >= 6 and len(R2) >= 3


@pytest.mark.parametrize("name", R1)
@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_r1_image_matches_reference(name, dt):
    g = load_golden(name)
    rgb, alpha, depth = _render(g, dt)
    assert np.abs(rgb.numpy() - g["rgb"]).max() <= 2e-6
    assert np.abs(alpha.numpy() - g["alpha"]).max() <= 2e-6
    # depth = D/(W+1e-6) is ill-conditioned where W is tiny: compare where W matters
    m = g["alpha"] > 1e-3
    assert np.abs(depth.numpy() - g["depth"])[m].max() <= 1e-4
    assert rel_l2(depth.numpy(), g["depth"]) <= 1e-4


@pytest.mark.parametrize("name", R1)
@pytest.mark.parametrize("tag", ["nodepth", "depth"])
def test_r1_gradients_match_reference_autograd(name, tag):
    g = load_golden(name)
    dt = torch.float64
    leaves = [_t(g[k], dt).requires_grad_(True) for k in ("means", "scales", "colors", "opac")]
    rgb, alpha, depth = r1.render_r1(*leaves, _t(g["view"], dt), _t(g["proj"], dt), int(g["width"]),
                                     int(g["height"]), background=_t(g["bg"], dt))
    loss = (rgb * _t(g["g_rgb"], dt)).sum() + (alpha * _t(g["g_alpha"], dt)).sum()
    if tag == "depth":
        loss = loss + (depth * _t(g["g_depth"], dt)).sum()
    loss.backward()
    for leaf, key in zip(leaves, ("means", "scales", "colors", "opac")):
        assert rel_l2(leaf.grad.numpy(), g[f"grad_{key}_{tag}"]) <= 2e-4, key


@pytest.mark.parametrize("name", R1)
def test_c_projection_matches_reference_project(name):
    g = load_golden(name)
    W, H = int(g["width"]), int(g["height"])
    b = ocpu.bin_gaussians(g["means"], g["scales"], g["opac"], g["view"], g["proj"], W, H, k=5.0)
    v = g["valid"]
    assert np.allclose(b["px"][v], g["px"][v], rtol=1e-5, atol=1e-3)
    assert np.allclose(b["py"][v], g["py"][v], rtol=1e-5, atol=1e-3)
    assert np.allclose(b["zabs"], g["z_abs"], rtol=1e-6, atol=1e-7)
    # a culled Gaussian is either invalid in the reference, has op<=0, or is off-screen
    culled = b["cnt"] == 0
    assert np.all(~culled[(v & (g["opac"] > 0) & (b["px"] > 0) & (b["px"] < W - 1) & (b["py"] > 0) & (b["py"] < H - 1))])
    assert np.all(culled[~v])


@pytest.mark.parametrize("name", R1)
@pytest.mark.parametrize("begin_bit", [0, 32])
def test_c_binning_invariants(name, begin_bit):
    g = load_golden(name)
    W, H = int(g["width"]), int(g["height"])
    b = ocpu.bin_gaussians(g["means"], g["scales"], g["opac"], g["view"], g["proj"], W, H, k=5.0,
                           begin_bit=begin_bit)
    keys, vals, ranges = b["keys"], b["vals"], b["ranges"]
    assert b["total"] == int(b["cnt"].sum()) == len(keys)
    shifted = keys >> np.uint64(begin_bit)
    assert np.all(shifted[1:] >= shifted[:-1])                      # sortedness
    order = np.argsort(b["keys_unsorted"] >> np.uint64(begin_bit), kind="stable")
    assert np.array_equal(keys, b["keys_unsorted"][order])          # stability
    assert np.array_equal(vals, b["vals_unsorted"][order])
    tiles = (keys >> np.uint64(32)).astype(np.int64)
    for t in range(ranges.shape[0]):
        s, e = ranges[t]
        assert np.all(tiles[s:e] == t) and (e - s) == int((tiles == t).sum())
    # elliptical tile culling (torch style): every bbox pixel INSIDE the 5-sigma ellipse lies in a binned tile ...
    for i in np.nonzero(b["cnt"])[0][:50]:
        x0, y0, x1, y1 = b["bbox"][i]
        tl = set(tiles[vals == i].tolist())
        ys, xs = np.mgrid[y0:y1 + 1, x0:x1 + 1]
        q = ((xs + 0.5 - b["px"][i]) / b["sx"][i]) ** 2 + ((ys + 0.5 - b["py"][i]) / b["sy"][i]) ** 2
        inside = q <= 24.99
        assert set(((ys[inside] // 16) * b["tiles_x"] + xs[inside] // 16).tolist()) <= tl
    # ... and without culling every pixel of the bbox does
    u = ocpu.bin_gaussians(g["means"], g["scales"], g["opac"], g["view"], g["proj"], W, H, k=5.0,
                           begin_bit=begin_bit, cull=False)
    ut = (u["keys"] >> np.uint64(32)).astype(np.int64)
    assert u["total"] >= b["total"]
    for i in np.nonzero(u["cnt"])[0][:50]:
        x0, y0, x1, y1 = u["bbox"][i]
        tl = set(ut[u["vals"] == i].tolist())
        assert {(y // 16) * u["tiles_x"] + (x // 16) for y in (y0, y1) for x in (x0, x1)} <= tl
    # bbox formula of the reference renderers: max(0,floor(p-k*s)) .. min(W-1,ceil(p+k*s))
    on = b["cnt"] > 0
    assert np.array_equal(b["bbox"][on, 0], np.maximum(0, np.floor(b["px"] - 5 * b["sx"]))[on].astype(np.int32))
    assert np.array_equal(b["bbox"][on, 3], np.minimum(H - 1, np.ceil(b["py"] + 5 * b["sy"]))[on].astype(np.int32))


def test_depth_bits_order():
    zs = np.array([-5.0, -2.5, -2.4999, -1e-3, -0.0, 0.0, 1e-3, 3.0], np.float32)
    bits = [ocpu.b2o().b2o_depth_bits(float(z)) for z in zs]
    # larger camera z (closer, the reference sorts z descending) => smaller key
    assert all(bits[i] >= bits[i + 1] for i in range(len(bits) - 1))
    assert bits[0] > bits[1] > bits[2]


@pytest.mark.parametrize("name", R1)
def test_cutoff5_restatement_within_image_tolerance(name):
    """bbox cutoff at k=5 keeps RGB/alpha within 1e-4 of the un-truncated reference."""
    g = load_golden(name)
    rgb, alpha, _ = _render(g, torch.float32, cutoff_sigma=5.0, tile=16)
    assert np.abs(rgb.numpy() - g["rgb"]).max() <= 1e-4
    assert np.abs(alpha.numpy() - g["alpha"]).max() <= 1e-4


@pytest.mark.parametrize("name", R1)
def test_c_blend_over_bins_matches_reference(name):
    g = load_golden(name)
    W, H = int(g["width"]), int(g["height"])
    b = ocpu.bin_gaussians(g["means"], g["scales"], g["opac"], g["view"], g["proj"], W, H, k=5.0, begin_bit=32)
    col = r1.eval_colors(_t(g["colors"], torch.float32), _t(g["means"], torch.float32),
                         _t(g["view"], torch.float32)).numpy()
    rgb, alpha, depth = ocpu.blend_wsum(b, g["opac"], col, W, H, g["bg"])
    assert np.abs(rgb - g["rgb"]).max() <= 1e-4
    assert np.abs(alpha - g["alpha"]).max() <= 1e-4


@pytest.mark.parametrize("name", R2)
def test_sorted_restatement_matches_reference_cpu_renderer(name):
    g = load_golden(name)
    dt = torch.float32
    args = [_t(g[k], dt) for k in ("means", "scales", "colors", "opac", "view", "proj")]
    rgb, _ = r1.render_sorted(*args, int(g["width"]), int(g["height"]), background=_t(g["bg"], dt))
    q = r1.quantise_rgba8(rgb).numpy().astype(np.int32)
    ref = g["rgba_sorted"].astype(np.int32)
    assert np.abs(q - ref).max() <= 1
    assert (q == ref).mean() > 0.995


@pytest.mark.skipif(not ocpu.have_r2ref(), reason="oracle/_ref/libr2ref.so not built")
@pytest.mark.parametrize("name", R2)
def test_compiled_reference_reproduces_golden(name):
    g = load_golden(name)
    img = ocpu.r2_render(g["means"], g["scales"], g["colors"], g["opac"], g["view"], g["proj"],
                         int(g["width"]), int(g["height"]), g["bg"], depth_sort=1)
    assert np.array_equal(img, g["rgba_sorted"])


def test_adam_restatement_matches_torch():
    torch.manual_seed(0)
    p0 = torch.randn(257, dtype=torch.float32)
    p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([p], lr=0.02)
    q, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    for step in range(1, 6):
        g = torch.randn(257)
        p.grad = g.clone()
        opt.step()
        q, m, v = r1.adam_step(q, g, m, v, step, 0.02)
        assert torch.allclose(q, p.detach(), rtol=1e-5, atol=1e-7)


# ---------------------------------------------------------------- densify / prune ----------
@pytest.mark.parametrize("name", golden_names("densify_"))
def test_densify_oracle_matches_reference_function(name):
    """oracle/fit_oracle.py vs the unmodified _densify_and_prune (fit_multiview_stub.py:140-197),
    including the N(0,1) jitter (same CPU generator seed and call order)."""
    from oracle import fit_oracle
    g = load_golden(name)
    t = lambda k: torch.from_numpy(g[k])
    torch.manual_seed(int(g["seed"]))
    m, s, o, c, keep, src = fit_oracle.densify_and_prune(t("means"), t("scales_raw"), t("op_raw"), t("colors"),
                                                         int(g["max_gaussians"]), float(g["ratio"]),
                                                         float(g["prune_opacity"]))
    assert np.array_equal(m.numpy(), g["out_means"])
    assert np.array_equal(s.numpy(), g["out_scales_raw"])
    assert np.array_equal(o.numpy(), g["out_op_raw"])
    assert np.array_equal(c.numpy(), g["out_colors"])
    assert m.shape[0] == int(keep.sum()) + src.shape[0]


@pytest.mark.parametrize("name", R1)
def test_r1_pixel_subset_is_the_full_render_on_those_pixels(name):
    """render_r1(pixels=...) -- the crop oracle of the full-size GPU parity test -- reproduces the reference's values
    and gradients on the chosen pixels (cotangent zero elsewhere)."""
    g = load_golden(name)
    W, H = int(g["width"]), int(g["height"])
    rng = np.random.RandomState(0)
    k = min(50, W * H)
    flat = rng.choice(W * H, size=k, replace=False)
    ix, iy = torch.from_numpy(flat % W), torch.from_numpy(flat // W)
    t = lambda a: torch.from_numpy(a).to(torch.float64)
    leaves = [t(g[key]).requires_grad_(True) for key in ("means", "scales", "colors", "opac")]
    rgb, alpha, depth = r1.render_r1(*leaves, t(g["view"]), t(g["proj"]), W, H, background=t(g["bg"]), pixels=(ix, iy))
    assert np.abs(rgb.detach().numpy() - g["rgb"][iy, ix]).max() <= 1e-5
    assert np.abs(alpha.detach().numpy() - g["alpha"][iy, ix]).max() <= 1e-5
    cot = torch.from_numpy(g["g_rgb"][iy, ix]).to(torch.float64)
    (rgb * cot).sum().backward()
    full = [t(g[key]).requires_grad_(True) for key in ("means", "scales", "colors", "opac")]
    rgb_f, _, _ = r1.render_r1(*full, t(g["view"]), t(g["proj"]), W, H, background=t(g["bg"]))
    (rgb_f[iy, ix] * cot).sum().backward()
    for a, b in zip(leaves, full):
        assert rel_l2(a.grad.numpy(), b.grad.numpy()) <= 1e-12
