"""C3 viewer config (1 M Gaussians, 960x540, depth sorted), a few frames: target for ncu launch lists."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
r = importlib.import_module("3dgaussian_b200.renderer")
capi = importlib.import_module("3dgaussian_b200.capi")
synth = importlib.import_module("3dgaussian_b200.synth")
dev = torch.device("cuda", 0)
n, W, H = 1_000_000, 960, 540
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ds = int(sys.argv[2]) if len(sys.argv) > 2 else 1
m, s, c, o = synth.synth_gaussians(n, 1, 1234, dev, 0.004, 0.02)
view, proj = synth.orbit_camera(0, 1, W, H)
params = capi.make_params(W, H, view.reshape(-1).tolist(), proj.reshape(-1).tolist(), (0.02, 0.02, 0.02),
                          mode=capi.MODE_SORTED if ds else capi.MODE_WSUM, style=capi.STYLE_NATIVE, cutoff_sigma=3.0,
                          sh_coeffs=1, sort_depth=ds, exact_bbox=1)
mp = int(r.count_pairs(params, m, s, o) * 1.25) + 4096
L = capi.lib()
ws = torch.empty(L.b2s_workspace_bytes(n, W, H, mp) + L.b2s_state_bytes(n, W, H, mp), dtype=torch.uint8, device=dev)
img = torch.empty((H, W, 4), dtype=torch.uint8, device=dev)
for _ in range(frames):
    r.render_rgba8(m, s, c, o, view, proj, W, H, (0.02, 0.02, 0.02), enable_depth_sort=ds, max_pairs=mp, out=img, workspace=ws)
torch.cuda.synchronize()
print("frames", frames, "checksum", int(img.sum()))
if os.environ.get("B2S_STATS_DUMP"):
    wb = L.b2s_workspace_bytes(n, W, H, mp)
    c = ws[wb:wb + 32].view(torch.int32).cpu().tolist()
    print("counters", c, "-> REST chunks", c[5], "REST units run", c[6], "FIRST tiles saturated", c[7], "(accumulated over", frames, "frames: pad is never cleared)")
