import faulthandler, importlib, os, sys
faulthandler.enable()
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, ROOT)
import numpy as np, torch
import scenes
r = importlib.import_module("3dgaussian_b200.renderer")
means, scales, colors, opac = scenes.make_scene(11, 3000, sh=1, s_lo=0.01, s_hi=0.05)
view, proj = scenes.orbit_camera(0, 4, 160, 90)
bg = np.array([0.02, 0.02, 0.02], np.float32)
which = sys.argv[1] if len(sys.argv) > 1 else "host"
if which == "dev_first":
    d = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(a).to(d)
    print("device entry", flush=True)
    r.render_rgba8(t(means), t(scales), t(colors), t(opac), view, proj, 160, 90, bg)
    torch.cuda.synchronize()
print("host entry", flush=True)
img = r.render_gaussians(means, scales, colors, opac, 160, 90, view, proj, bg, enable_depth_sort=1)
print("ok", img.shape, img.mean(), flush=True)
