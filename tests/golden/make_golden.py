"""Generates tests/golden/*.npz by running the UNMODIFIED reference:
  - R1: /root/reference/python/torch_renderer.py (imported by path), CPU float32,
        outputs + autograd gradients for fixed random cotangents
  - R2: /root/reference/src/renderer_cpu.cpp via oracle/_ref/libr2ref.so (RGBA8)
Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The .npz files are committed; this script is the provenance record.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes  # noqa: E402
from oracle import cpu as ocpu  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_torch_renderer", "/root/reference/python/torch_renderer.py")
ref = importlib.util.module_from_spec(spec)
sys.modules["ref_torch_renderer"] = ref   # dataclass machinery looks the module up by name
spec.loader.exec_module(ref)

torch.manual_seed(0)
torch.set_num_threads(4)

R1_CASES = [
    # name, seed, n, sh, W, H, bg, edge, s_lo, s_hi, view index/num
    ("r1_rgb_basic", 11, 96, 1, 40, 32, (0.0, 0.0, 0.0), False, 0.02, 0.2, (0, 4)),
    ("r1_sh4_bg", 12, 96, 4, 40, 32, (0.1, 0.2, 0.3), False, 0.02, 0.2, (1, 4)),
    ("r1_edge_rgb", 13, 64, 1, 33, 25, (0.05, 0.0, 0.9), True, 0.02, 0.3, (2, 4)),
    ("r1_edge_sh4", 14, 64, 4, 33, 25, (0.0, 0.0, 0.0), True, 0.02, 0.3, (3, 4)),
    ("r1_many_small", 15, 400, 1, 64, 48, (0.0, 0.0, 0.0), False, 0.005, 0.05, (1, 8)),
    ("r1_fit_like", 16, 300, 4, 48, 48, (0.0, 0.0, 0.0), False, 0.09, 0.11, (0, 4)),
]


def run_r1(name, seed, n, sh, W, H, bg, edge, s_lo, s_hi, vi):
    means, scales, colors, opac = scenes.make_scene(seed, n, sh=sh, s_lo=s_lo, s_hi=s_hi, edge_cases=edge)
    view, proj = scenes.orbit_camera(vi[0], vi[1], W, H)
    r = np.random.RandomState(seed + 1000)
    g_rgb = r.randn(H, W, 3).astype(np.float32)
    g_alpha = r.randn(H, W).astype(np.float32)
    g_depth = (0.1 * r.randn(H, W)).astype(np.float32)
    t = lambda a: torch.from_numpy(a.copy()).requires_grad_(True)
    out = {}
    for tag, use_depth in (("nodepth", False), ("depth", True)):
        tm, ts, tc, to = t(means), t(scales), t(colors), t(opac)
        cam = ref.Camera(view=torch.from_numpy(view), proj=torch.from_numpy(proj))
        rgb, alpha, depth = ref.render_gaussians_torch(
            tm, ts, tc, to, cam, width=W, height=H, background=torch.tensor(bg, dtype=torch.float32),
            max_gaussians=100000, return_aux=True)
        loss = (rgb * torch.from_numpy(g_rgb)).sum() + (alpha * torch.from_numpy(g_alpha)).sum()
        if use_depth:
            loss = loss + (depth * torch.from_numpy(g_depth)).sum()
        loss.backward()
        out[f"grad_means_{tag}"] = tm.grad.numpy()
        out[f"grad_scales_{tag}"] = ts.grad.numpy()
        out[f"grad_colors_{tag}"] = tc.grad.numpy()
        out[f"grad_opac_{tag}"] = to.grad.numpy()
    # projection intermediates straight from the reference's _project
    px, py, z_abs, valid = ref._project(torch.from_numpy(means), torch.from_numpy(view), torch.from_numpy(proj), W, H)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        means=means, scales=scales, colors=colors, opac=opac, view=view, proj=proj,
        bg=np.asarray(bg, np.float32), width=W, height=H,
        g_rgb=g_rgb, g_alpha=g_alpha, g_depth=g_depth,
        rgb=rgb.detach().numpy(), alpha=alpha.detach().numpy(), depth=depth.detach().numpy(),
        px=px.numpy(), py=py.numpy(), z_abs=z_abs.numpy(), valid=valid.numpy(), **out)
    print("wrote", name, "rgb mean", float(rgb.mean()))


R2_CASES = [
    # name, seed, n, W, H, bg, s_lo, s_hi, edge
    ("r2_viewer_small", 21, 300, 64, 48, (0.02, 0.02, 0.02), 0.02, 0.15, False),
    ("r2_viewer_edge", 22, 128, 45, 35, (0.02, 0.02, 0.02), 0.02, 0.3, True),
    ("r2_dense", 23, 1500, 96, 54, (0.02, 0.02, 0.02), 0.01, 0.05, False),
]


def run_r2(name, seed, n, W, H, bg, s_lo, s_hi, edge):
    means, scales, colors, opac = scenes.make_scene(seed, n, sh=1, s_lo=s_lo, s_hi=s_hi, edge_cases=edge)
    colors = np.clip(colors, 0, 1)  # the native path takes displayable colours
    view, proj = scenes.orbit_camera(0, 1, W, H)   # viewer start pose: yaw 0, pitch 0.2, r 2.5
    sorted_img = ocpu.r2_render(means, scales, colors, opac, view, proj, W, H, bg, depth_sort=1)
    wsum_img = ocpu.r2_render(means, scales, colors, opac, view, proj, W, H, bg, depth_sort=0)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), means=means, scales=scales, colors=colors,
                        opac=opac, view=view, proj=proj, bg=np.asarray(bg, np.float32), width=W, height=H,
                        rgba_sorted=sorted_img, rgba_wsum=wsum_img)
    print("wrote", name, "mean", sorted_img[..., :3].mean(), wsum_img[..., :3].mean())


DENSIFY_CASES = [
    # name, seed, n, sh, max_gaussians, ratio, prune_opacity, op_shift
    ("densify_basic", 31, 500, 1, 3000, 0.15, 0.05, -2.0),
    ("densify_room_limited", 32, 400, 4, 420, 0.15, 0.05, -1.0),
    ("densify_top64", 33, 300, 1, 3000, 0.15, 0.05, -6.0),      # almost everything below the threshold
]


def run_densify(name, seed, n, sh, max_g, ratio, thr, op_shift):
    """The unmodified reference _densify_and_prune (fit_multiview_stub.py:140-197) on seeded raw parameters."""
    spec2 = importlib.util.spec_from_file_location("ref_fit", "/root/reference/python/fit_multiview_stub.py")
    sys.path.insert(0, "/root/reference/python")
    fitmod = importlib.util.module_from_spec(spec2)
    spec2.loader.exec_module(fitmod)
    sys.path.pop(0)
    r = np.random.RandomState(seed)
    means = ((r.rand(n, 3) - 0.5) * 1.2).astype(np.float32)
    scales_raw = (-2.2 + 0.5 * r.randn(n, 3)).astype(np.float32)
    op_raw = (op_shift + 1.5 * r.randn(n)).astype(np.float32)
    colors = (0.1 * r.rand(n, 3)).astype(np.float32) if sh == 1 else (0.1 * r.randn(n, sh, 3)).astype(np.float32)
    params = {"means": torch.nn.Parameter(torch.from_numpy(means.copy())),
              "scales_raw": torch.nn.Parameter(torch.from_numpy(scales_raw.copy())),
              "opacities_raw": torch.nn.Parameter(torch.from_numpy(op_raw.copy()))}
    params["colors_raw" if sh == 1 else "sh_raw"] = torch.nn.Parameter(torch.from_numpy(colors.copy()))
    torch.manual_seed(seed)
    out = fitmod._densify_and_prune(params, max_g, ratio, thr)
    oc = out["colors_raw" if sh == 1 else "sh_raw"]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), means=means, scales_raw=scales_raw, op_raw=op_raw,
                        colors=colors, max_gaussians=max_g, ratio=ratio, prune_opacity=thr, seed=seed,
                        out_means=out["means"].detach().numpy(), out_scales_raw=out["scales_raw"].detach().numpy(),
                        out_op_raw=out["opacities_raw"].detach().numpy(), out_colors=oc.detach().numpy())
    print("wrote", name, "n", n, "->", out["means"].shape[0])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "densify":
        for c in DENSIFY_CASES:
            run_densify(*c)
        sys.exit(0)
    ocpu.build()
    for c in R1_CASES:
        run_r1(*c)
    for c in R2_CASES:
        run_r2(*c)
    for c in DENSIFY_CASES:
        run_densify(*c)
