"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck / initcheck / synccheck): every kernel family
of the hot path once -- drop-in forward + backward without and with the depth plane (tcgen05 kernels, multi-unit
tiles), a FitDriver iteration with mask + depth terms on two lanes, guarded Adam, and a depth-sorted RGBA8 frame."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch

r = importlib.import_module("3dgaussian_b200.renderer")
fit = importlib.import_module("3dgaussian_b200.fit")
synth = importlib.import_module("3dgaussian_b200.synth")
capi = importlib.import_module("3dgaussian_b200.capi")
dev = torch.device("cuda", 0)
n, W, H = 1500, 64, 48
m, s, c, o = synth.synth_gaussians(n, 4, 7, dev, 0.05, 0.3)        # large splats: > 512 Gaussians per tile (multi-unit)
view, proj = synth.orbit_camera(1, 4, W, H)
cam = r.Camera(view=torch.from_numpy(view).to(dev), proj=torch.from_numpy(proj).to(dev))
leaves = [x.clone().requires_grad_(True) for x in (m, s, c, o)]
img = r.render_gaussians_torch(*leaves, cam, W, H, max_gaussians=n)
img.square().sum().backward()
leaves = [x.clone().requires_grad_(True) for x in (m, s, c, o)]
rgb, alpha, depth = r.render_gaussians_torch(*leaves, cam, W, H, max_gaussians=n, return_aux=True,
                                             background=torch.tensor([0.1, 0.2, 0.3], device=dev))
(rgb.sum() + alpha.sum() + 0.1 * depth.sum()).backward()
torch.cuda.synchronize()
cams = synth.orbit_cameras(3, W, H)
d = fit.FitDriver(n, 4, W, H, cams, dev, lanes=2, use_depth=True)
sr, orr, cr = synth.to_raw(s, o, c, 4)
d.set_params(m, sr, orr, cr)
d.plan()
g = torch.Generator(device=dev).manual_seed(0)
d.set_targets({i: torch.rand((H, W, 3), generator=g, device=dev) for i in d.views},
              {i: (torch.rand((H, W), generator=g, device=dev) > 0.5).float() for i in d.views},
              {i: torch.rand((H, W), generator=g, device=dev) for i in d.views})
for _ in range(2):
    d.step()
assert not d.check_overflow()
mm, ss, cc, oo = synth.synth_gaussians(3000, 1, 3, dev, 0.01, 0.05)
im8 = r.render_rgba8(mm, ss, cc, oo, view, proj, W, H, (0.02, 0.02, 0.02), enable_depth_sort=1)
im8b = r.render_rgba8(mm, ss, cc, oo, view, proj, W, H, (0.02, 0.02, 0.02), enable_depth_sort=0)
torch.cuda.synchronize()
print("SANITIZE_TARGET_OK", capi.path_counts(), float(img.mean()), float(d.loss_dev.item()), int(im8.sum()), flush=True)
