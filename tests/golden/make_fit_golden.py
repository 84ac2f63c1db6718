"""Generates the end-to-end fixtures of the fit configs (BASELINE.json configs[0..1]) by running the UNMODIFIED
reference script /root/reference/python/fit_multiview_stub.py on the CPU with its own torch_renderer.py
(torch.manual_seed(0), CPU generator) in the build container:

    tests/golden/c1_cpu_loss.txt        config 1: --iters 150 --width 128 --height 128, 4 views of assets/scene_tex
    tests/golden/c2_inputs/{masks,depth}/NN.png   config 2 inputs the reference does not ship (README.md:43-44 points
                                        at a non-existent data/ dir): silhouette masks = mean RGB > 0.06 of the resized
                                        targets (the script's own estimate, :37-42), depth maps = normalised R1 depth
                                        of a seeded 300-Gaussian blob set seen from the script's orbit cameras
    tests/golden/c2_cpu_loss.txt        config 2 (256x256 --use_sh, masks + depth, 1200 Gaussians), first 4 iterations
                                        (one iteration takes ~40 s of CPU)

    python tests/golden/make_fit_golden.py [c1] [c2inputs] [c2]
"""
import os
import runpy
import shutil
import sys
import tempfile

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
C2_ARGS = ["--width", "256", "--height", "256", "--use_sh", "--num_gaussians", "1200", "--max_gaussians", "3000",
           "--densify_interval", "40", "--prune_interval", "40"]


def run_reference_cpu(args, seed=0):
    out = tempfile.mkdtemp()
    sys.path.insert(0, os.path.join(REF, "python"))
    for name in ("torch_renderer", "device_utils"):
        sys.modules.pop(name, None)
    torch.manual_seed(seed)
    argv = sys.argv
    sys.argv = ["fit_multiview_stub.py", "--targets_dir", os.path.join(REF, "assets", "scene_tex"), "--out_dir", out] + args
    try:
        runpy.run_path(os.path.join(REF, "python", "fit_multiview_stub.py"), run_name="__main__")
    finally:
        sys.argv = argv
        sys.path.pop(0)
    loss = open(os.path.join(out, "loss.txt")).read()
    shutil.rmtree(out)
    return loss


def make_c2_inputs():
    sys.path.insert(0, os.path.join(REF, "python"))
    import torch_renderer as tr
    root = os.path.join(HERE, "c2_inputs")
    os.makedirs(os.path.join(root, "masks"), exist_ok=True)
    os.makedirs(os.path.join(root, "depth"), exist_ok=True)
    W = H = 256
    g = torch.Generator().manual_seed(4321)
    n = 300
    means = (torch.rand((n, 3), generator=g) - 0.5) * 0.9
    scales = torch.full((n, 3), 0.09)
    colors = torch.rand((n, 3), generator=g)
    opac = torch.full((n,), 0.8)
    names = sorted(os.listdir(os.path.join(REF, "assets", "scene_tex")))
    proj = tr.perspective(60.0, 1.0, 0.01, 100.0)
    import math
    for i, nm in enumerate(names):
        stem = os.path.splitext(nm)[0]
        img = Image.open(os.path.join(REF, "assets", "scene_tex", nm)).convert("RGB").resize((W, H), Image.Resampling.BILINEAR)
        t = np.asarray(img, dtype=np.float32) / 255.0
        mask = (t.mean(axis=2) > 0.06).astype(np.uint8) * 255
        Image.fromarray(mask, mode="L").save(os.path.join(root, "masks", stem + ".png"))
        yaw = 2.0 * math.pi * i / len(names)
        eye = torch.tensor([2.5 * math.cos(0.2) * math.sin(yaw), 2.5 * math.sin(0.2), 2.5 * math.cos(0.2) * math.cos(yaw)])
        view = tr.look_at(eye, torch.zeros(3), torch.tensor([0.0, 1.0, 0.0]))
        _, alpha, depth = tr.render_gaussians_torch(means, scales, colors, opac, tr.Camera(view=view, proj=proj), W, H,
                                                    return_aux=True)
        d = depth * (alpha > 0.05)
        d8 = (d / (d.max() + 1e-6) * 255.0 + 0.5).clamp(0, 255).to(torch.uint8).numpy()
        Image.fromarray(d8, mode="L").save(os.path.join(root, "depth", stem + ".png"))
    sys.path.pop(0)


if __name__ == "__main__":
    what = sys.argv[1:] or ["c1", "c2inputs", "c2"]
    sys.argv = sys.argv[:1]
    if "c2inputs" in what:
        make_c2_inputs()
    if "c1" in what:
        open(os.path.join(HERE, "c1_cpu_loss.txt"), "w").write(run_reference_cpu(["--iters", "150", "--width", "128", "--height", "128"]))
    if "c2" in what:
        root = os.path.join(HERE, "c2_inputs")
        open(os.path.join(HERE, "c2_cpu_loss.txt"), "w").write(
            run_reference_cpu(["--iters", "4", "--masks_dir", os.path.join(root, "masks"), "--depth_dir", os.path.join(root, "depth")] + C2_ARGS))
