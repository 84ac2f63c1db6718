"""Step-by-step exercise of the C ABI with a flushed print before every native call (crash localisation)."""
import faulthandler
import importlib
import os
import sys

faulthandler.enable()
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch

def say(*a):
    print(*a, flush=True)

capi = importlib.import_module("3dgaussian_b200.capi")
r = importlib.import_module("3dgaussian_b200.renderer")
synth = importlib.import_module("3dgaussian_b200.synth")
say("lib", capi.lib().b2s_version())
say("sm_count", capi.lib().b2s_sm_count())
say("paths", capi.path_counts())
dev = torch.device("cuda", 0)
m, s, c, o = synth.synth_gaussians(2000, 1, 1, dev, 0.02, 0.1)
view, proj = synth.orbit_camera(0, 4, 96, 64)
cam = r.Camera(view=torch.from_numpy(view).to(dev), proj=torch.from_numpy(proj).to(dev))
say("ctx", capi.ctx(0))
p = capi.make_params(96, 64, view.reshape(-1), proj.reshape(-1))
say("count_pairs")
say(r.count_pairs(p, m, s, o))
say("forward no aux")
m.requires_grad_(True)
img = r.render_gaussians_torch(m, s, c, o, cam, 96, 64)
torch.cuda.synchronize()
say("ticket", capi.ticket_info(0))
say("backward")
img.sum().backward()
torch.cuda.synchronize()
say("forward aux + device background")
out = r.render_gaussians_torch(m, s, c, o, cam, 96, 64, background=torch.tensor([0.1, 0.2, 0.3], device=dev), return_aux=True)
torch.cuda.synchronize()
say("rgba8")
im8 = r.render_rgba8(m.detach(), s, c, o, view, proj, 96, 64)
torch.cuda.synchronize()
say("fit driver")
fit = importlib.import_module("3dgaussian_b200.fit")
cams = synth.orbit_cameras(4, 96, 64)
d = fit.FitDriver(2000, 1, 96, 64, cams, dev, lanes=2)
sr, orr, cr = synth.to_raw(s, o, c, 1)
d.set_params(m.detach(), sr, orr, cr)
say("plan")
d.plan()
say("render_view")
t = {i: d.render_view(i)[0].clone() for i in d.views}
d.set_targets(t, {})
say("step")
d.step()
torch.cuda.synchronize()
say("paths", capi.path_counts())
say("DIAG_OK")
