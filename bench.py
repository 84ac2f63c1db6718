#!/usr/bin/env python
"""Benchmark of the B200-native Gaussian-splatting fit iteration (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (rank 0 only)

One step = one fit iteration = forward + backward over ALL views + Adam
(reference python/fit_multiview_stub.py:265-311) on the synthetic workload of BASELINE.json
configs[3]: 1 M Gaussians, SH degree 3 (N,16,3), 64 orbit views at 1920x1080, views sharded
round-robin over the ranks, one NCCL all-reduce of the flat gradient buffer per iteration.
Prints ONE JSON line on rank 0 (see the task contract for the keys).
"""
from __future__ import annotations

import argparse
import importlib
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

synth = importlib.import_module("3dgaussian_b200.synth")     # workload generators (package side, no test-tree import)
synth_gaussians, to_raw, cameras = synth.synth_gaussians, synth.to_raw, synth.orbit_cameras

METRIC = "fit iters/s (fwd+bwd+Adam, all views)"
LANES_PER_SM = 128       # FP32 lanes per SM; the SM count comes from the device (b2s_sm_count)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    # workload overrides (the defaults ARE the benchmark; overrides are for smoke tests)
    ap.add_argument("--gaussians", "--n", dest="n", type=int, default=1_000_000)
    ap.add_argument("--sh", type=int, default=16)
    ap.add_argument("--views", type=int, default=64)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--lanes", type=int, default=4, help="concurrent view lanes (CUDA streams) per GPU")
    ap.add_argument("--view-groups", type=int, default=0,
                    help="pipeline the local views in this many groups (preprocess ahead, chain rule behind); 0 = auto")
    ap.add_argument("--grad-chunks", type=int, default=0,
                    help="Gaussian chunks of the pipelined multi-GPU tail (chain rule | all-reduce | Adam); 0 = FitDriver default")
    ap.add_argument("--comm", default="", choices=["", "nccl", "multimem"],
                    help="multi-GPU tail: fused NVLink-multicast reduce-scatter + Adam + all-gather kernel (multimem, the "
                         "default when available) or NCCL all-reduce + replicated Adam")
    ap.add_argument("--timeline", action="store_true",
                    help="record CUDA events at the phase boundaries of every timed step on rank 0 (head / views / tail)")
    ap.add_argument("--no-reorder", action="store_true", help="keep the synthetic Gaussians in generation (random) order")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-render", action="store_true")
    ap.add_argument("--no-timing", action="store_true", help="do not bracket stages with CUDA events")
    ap.add_argument("--s-lo", type=float, default=0.004, help="synthetic scale range (SURVEY 8d: 0.004..0.02 at 1 M)")
    ap.add_argument("--s-hi", type=float, default=0.02)
    ap.add_argument("--densify-every", type=int, default=0,
                    help="call FitDriver.densify_prune every K timed steps (BASELINE configs[4]); 0 = never")
    ap.add_argument("--render-only", action="store_true", help="only the render config (BASELINE configs[2]); for profiling")
    ap.add_argument("--cutoff-sigma", type=float, default=0.0,
                    help="bbox radius of the fit in sigmas (default: FitDriver's 5; the reference has none). For the record only: "
                         "the headline is measured at the default")
    ap.add_argument("--ext", action="store_true",
                    help="only the extension modes (rotations + EWA covariance, differentiable front-to-back compositing): "
                         "forward and forward+backward time of one view through render_gaussians_torch")
    ap.add_argument("--fit-scripts", action="store_true",
                    help="only BASELINE configs[0] and [1]: the reference's UNCHANGED fit_multiview_stub.py on the B200 path, "
                         "beside the same script on the box's host cores (the reference's stock CPU configuration)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
def bbox_pairs(means, scales, view, proj, W, H, k):
    """Algorithmic pixel-pair count P2 = sum of clamped k-sigma bbox areas (statistic only)."""
    V = torch.tensor(view, device=means.device).view(4, 4)
    P = torch.tensor(proj, device=means.device).view(4, 4)
    hom = torch.cat([means, torch.ones_like(means[:, :1])], 1)
    cam = hom @ V.t()
    clip = cam @ P.t()
    w = clip[:, 3]
    ws = torch.where(w.abs() < 1e-8, torch.ones_like(w), w)
    ndc = clip[:, :3] / ws[:, None]
    px = (ndc[:, 0] * 0.5 + 0.5) * (W - 1)
    py = (1 - (ndc[:, 1] * 0.5 + 0.5)) * (H - 1)
    z = cam[:, 2].abs().clamp_min(1e-6)
    sx = (scales[:, 0].abs() * 0.5 * W * P[0, 0].abs() / z).clamp_min(1.0)
    sy = (scales[:, 1].abs() * 0.5 * H * P[1, 1].abs() / z).clamp_min(1.0)
    ok = (ndc[:, 2] >= -1) & (ndc[:, 2] <= 1) & (w != 0)
    x0 = torch.floor(px - k * sx).clamp_min(0); x1 = torch.ceil(px + k * sx).clamp_max(W - 1)
    y0 = torch.floor(py - k * sy).clamp_min(0); y1 = torch.ceil(py + k * sy).clamp_max(H - 1)
    area = (x1 - x0 + 1).clamp_min(0) * (y1 - y0 + 1).clamp_min(0)
    return float((area * ok).sum().item())


class NvmlSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region through NVML (what nvidia-smi reads), from a
    thread every ~2 ms: the timed region of an 8-GPU run is 30 ms, shorter than nvidia-smi's start-up and its 100 ms loop
    period, which left that run without a single sample.  Falls back to the nvidia-smi loop when NVML is unavailable."""
    REASONS = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}

    def __init__(self, torch_device_index):
        self.fallback = None
        self.all, self.stop_flag, self.thread, self.t0 = [], False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(torch_device_index).uuid)
            h = None
            for i in range(pynvml.nvmlDeviceGetCount()):
                cand = pynvml.nvmlDeviceGetHandleByIndex(i)
                u = pynvml.nvmlDeviceGetUUID(cand)
                u = u.decode() if isinstance(u, bytes) else u
                if uuid in u:
                    h = cand
                    break
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(torch_device_index)
            self.nv, self.h = pynvml, h
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None
            self.fallback = ClockSampler(torch_device_index)

    def _loop(self):
        nv, h = self.nv, self.h
        while not self.stop_flag:
            try:
                self.all.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)),
                                 nv.nvmlDeviceGetPowerUsage(h) / 1000.0,
                                 int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.fallback is not None:
            return None
        import threading
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def begin(self):
        """Start of the timed window (the nvidia-smi fallback starts its loop here)."""
        if self.fallback is not None:
            return self.fallback.start()
        self.t0 = time.perf_counter()

    def end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.fallback is not None:
            return self.fallback.stop()
        t1 = getattr(self, "t1", None) or time.perf_counter()
        self.stop_flag = True
        self.thread.join(timeout=2)
        t0 = self.t0 if self.t0 is not None else 0.0
        self.samples = [(c, p, r) for t, c, p, r in self.all if t0 <= t <= t1]
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"], "source": "nvml"}
        reasons = sorted(nm for nm, bit in self.REASONS.items() if any(r & bit for _, _, r in self.samples))
        return {"sm_mhz": float(np.median([c for c, _, _ in self.samples])), "sm_max_mhz": self.max_mhz,
                "power_w_max": max(p for _, p, _ in self.samples), "samples": len(self.samples), "reasons": reasons,
                "source": "nvml, ~2 ms period, timed region only"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                sm.append(float(c[1])); mx.append(float(c[2]))
                try:
                    pw.append(float(c[3]))
                except ValueError:
                    pass
                for nm, val in zip(names, c[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
def _ref_threads():
    """Host threads of the CPU arm, set explicitly (torchrun exports OMP_NUM_THREADS=1, which silently made the
    N>=2 reference runs single-threaded in round 1)."""
    n = int(os.environ.get("B2S_REF_THREADS", "0")) or min(32, os.cpu_count() or 1)
    torch.set_num_threads(n)
    return n


def _reference_renderer():
    """(render function, fit-loss function, kind).  The UNMODIFIED reference module python/torch_renderer.py when it
    is staged under oracle/_ref/reference (oracle/Makefile `stage`; it travels with the snapshot) -- kind
    "reference"; otherwise the oracle port of it (oracle/r1_oracle.py) -- kind "port"."""
    from oracle import r1_oracle as r1
    ref_py = os.path.join(ROOT, "oracle", "_ref", "reference", "python")
    if os.path.exists(os.path.join(ref_py, "torch_renderer.py")):
        import importlib.util
        spec = importlib.util.spec_from_file_location("_reference_torch_renderer", os.path.join(ref_py, "torch_renderer.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)

        def render(means, sc, col, op, tv, tp, w, h):
            return mod.render_gaussians_torch(means, sc, col, op, mod.Camera(view=tv, proj=tp), w, h,
                                              max_gaussians=max(10000, means.shape[0]), return_aux=True)
        return render, r1.fit_loss, "reference"
    return (lambda means, sc, col, op, tv, tp, w, h: r1.render_r1(means, sc, col, op, tv, tp, w, h, chunk=256)), r1.fit_loss, "port"


def cpu_fit_sample(args, n_slice=1024, w=480, h=270, steps=1, warmup=0):
    """The reference's CPU path for this metric: R1 (python/torch_renderer.py) forward + autograd backward + torch
    Adam with the fit script's loss (fit_multiview_stub.py:292-311) on a bounded slice of the same workload -- the
    first n_slice Gaussians of the seed-1234 set, one orbit view, 480x270 -- on the host threads of _ref_threads().
    R1 is O(N*H*W) (no culling, no tiles): the full iteration (1.3e14 pairs) is extrapolated from the measured
    (Gaussian,pixel) pair rate.  The reference module supports (N,3)/(N,4,3) colours only (:106), so the slice
    uses its first 4 SH coefficients when the workload has more."""
    render, fit_loss, kind = _reference_renderer()
    threads = _ref_threads()
    dev = torch.device("cpu")
    means, scales, colors, opac = synth_gaussians(n_slice, args.sh, 1234, dev)
    sh = args.sh
    if kind == "reference" and sh > 4:
        colors, sh = colors[:, :4, :].contiguous(), 4
    scales_raw, op_raw, col_raw = to_raw(scales, opac, colors, sh)
    params = [torch.nn.Parameter(t.clone()) for t in (means, scales_raw, op_raw, col_raw)]
    opt = torch.optim.Adam(params, lr=0.02)
    view, proj = synth.orbit_camera(0, args.views, w, h)
    tv, tp = torch.from_numpy(view), torch.from_numpy(proj)
    tgt = torch.rand((h, w, 3), generator=torch.Generator().manual_seed(4321))
    mask = (tgt.mean(dim=2) > 0.06).float()
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        sc = torch.nn.functional.softplus(params[1]) + 1e-3
        op = torch.sigmoid(params[2])
        col = torch.sigmoid(params[3]) if sh == 1 else params[3]
        rgb, alpha, depth = render(params[0], sc, col, op, tv, tp, w, h)
        loss = fit_loss(rgb, alpha, depth, tgt, mask, None) + 1e-3 * op.mean() + 1e-3 * sc.mean()
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    sec = float(np.mean(times))
    pairs = float(n_slice) * w * h
    rate = pairs / sec                                   # dense (Gaussian,pixel) pairs per second
    full_pairs = float(args.n) * args.width * args.height * args.views
    what = ("the unmodified reference module python/torch_renderer.py (staged copy, oracle/_ref/reference)" if kind == "reference"
            else "R1 port (oracle/r1_oracle.py)")
    return {"sec_per_sample": sec, "pairs_per_s": rate, "iters_per_s_extrapolated": rate / full_pairs, "kind": kind,
            "sample": f"{what}: fwd+bwd+Adam, first {n_slice} Gaussians (SH{sh}), 1 view, {w}x{h}, {threads} threads; "
                      f"EXTRAPOLATED to the full iteration by dense pair count N*H*W*V (R1 has no culling)",
            "cores": threads}


def run_reference(args, out_fd):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_fit_sample(args, steps=max(1, args.steps), warmup=max(0, min(args.warmup, 1)))
    val = r["iters_per_s_extrapolated"]
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "iters/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / val, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": val, "unit": "iters/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                         "sec_per_sample": r["sec_per_sample"], "dense_pairs_per_s": r["pairs_per_s"]},
        "e2e": {"value": val, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(out_fd, line)


def VIEW_GROUPS_AUTO(nv_local, lanes):
    """One group: pipelining the local views in groups (preprocess ahead, chain rule behind) measured slower at 8, 16 and
    64 views per rank (6.82 -> 7.20 ms at 8 views: the chain rule re-reads the coefficients once per group)."""
    return 1


def workload_config(args):
    return {"workload": f"synthetic fit: {args.n} Gaussians SH{args.sh} (N,{args.sh},3), {args.views} orbit views at "
                        f"{args.width}x{args.height}, fwd+bwd+Adam, views sharded over ranks (BASELINE configs[3])",
            "gaussians": args.n, "sh_coeffs": args.sh, "views": args.views, "width": args.width, "height": args.height,
            "cutoff_sigma": (args.cutoff_sigma if args.cutoff_sigma > 0 else 5.0), "loss": "L1 recon + 0.2*L1 silhouette + 1e-3 reg", "parallelism": "views round-robin over ranks",
            "l2_note": "per-step inputs (params 220 MB + targets/masks 2.1 GB at N=1) exceed the 126 MB L2"}


# ------------------------------------------------------------------------------------------
def render_bench(device, frames=20):
    """BASELINE configs[2]: render-only forward, 1 M Gaussians, 960x540, native viewer params
    (enable_depth_sort=1, bg 0.02; reference src/model_viewer_main.cpp:193-200)."""
    r = importlib.import_module("3dgaussian_b200.renderer")
    n, W, H = 1_000_000, 960, 540
    means, scales, colors, opac = synth_gaussians(n, 1, 1234, device, 0.004, 0.02)
    view, proj = synth.orbit_camera(0, 1, W, H)
    out = {}
    capi = importlib.import_module("3dgaussian_b200.capi")
    for name, ds in (("sorted", 1), ("wsum", 0)):
        # a viewer keeps the model resident and replays frames: size the pair buffers once (one counting pass,
        # 25 % slack), then reuse workspace and output
        params = capi.make_params(W, H, view.reshape(-1).tolist(), proj.reshape(-1).tolist(), (0.02, 0.02, 0.02),
                                  mode=capi.MODE_SORTED if ds else capi.MODE_WSUM, style=capi.STYLE_NATIVE,
                                  cutoff_sigma=3.0, sh_coeffs=1, sort_depth=ds, exact_bbox=1)
        mp = int(r.count_pairs(params, means, scales, opac) * 1.25) + 4096
        L = capi.lib()
        ws = torch.empty(L.b2s_workspace_bytes(n, W, H, mp) + L.b2s_state_bytes(n, W, H, mp), dtype=torch.uint8, device=device)
        img = torch.empty((H, W, 4), dtype=torch.uint8, device=device)
        kw = dict(enable_depth_sort=ds, max_pairs=mp, out=img, workspace=ws)
        for _ in range(3):
            r.render_rgba8(means, scales, colors, opac, view, proj, W, H, (0.02, 0.02, 0.02), **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(frames):
            r.render_rgba8(means, scales, colors, opac, view, proj, W, H, (0.02, 0.02, 0.02), **kw)
        e1.record()
        torch.cuda.synchronize()
        out[f"ms_per_frame_{name}_device_resident"] = e0.elapsed_time(e1) / frames
        out[f"tile_pairs_{name}"] = int((mp - 4096) / 1.25)
        del ws, img
    # per-stage CUDA-event spans of the sorted frame (one more pass of `frames` frames with the brackets on)
    try:
        params = capi.make_params(W, H, view.reshape(-1).tolist(), proj.reshape(-1).tolist(), (0.02, 0.02, 0.02),
                                  mode=capi.MODE_SORTED, style=capi.STYLE_NATIVE, cutoff_sigma=3.0, sh_coeffs=1,
                                  sort_depth=1, exact_bbox=1)
        p1s = int(r.count_pairs(params, means, scales, opac))
        mp = int(p1s * 1.25) + 4096
        L = capi.lib()
        ws = torch.empty(L.b2s_workspace_bytes(n, W, H, mp) + L.b2s_state_bytes(n, W, H, mp), dtype=torch.uint8, device=device)
        img = torch.empty((H, W, 4), dtype=torch.uint8, device=device)
        kw = dict(enable_depth_sort=1, max_pairs=mp, out=img, workspace=ws)
        r.render_rgba8(means, scales, colors, opac, view, proj, W, H, (0.02, 0.02, 0.02), **kw)
        torch.cuda.synchronize()
        capi.timing_enable(device.index, True)
        capi.timing_read(device.index)
        for _ in range(frames):
            r.render_rgba8(means, scales, colors, opac, view, proj, W, H, (0.02, 0.02, 0.02), **kw)
        st = capi.timing_read(device.index)
        capi.timing_enable(device.index, False)
        out["stages_ms_per_frame_sorted"] = {k: v[0] / frames for k, v in st.items() if v[1] > 0}
        # roofline of the frame's HBM-bound stage, the 64-bit (tile|depth) radix sort: SURVEY 8(d) B_sort = 24 B per
        # tile-pair per 8-bit digit pass (12 B read + 12 B written), passes = ceil((tile bits + 32)/8)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        n_tiles = ((W + 15) // 16) * ((H + 15) // 16)
        passes = (max(1, math.ceil(math.log2(max(n_tiles, 2)))) + 32 + 7) // 8
        sort_ms = out["stages_ms_per_frame_sorted"].get("sort")
        if sort_ms:
            gbs = p1s * 24.0 * passes / (sort_ms * 1e-3) / 1e9
            out["roofline"] = {"kernel": "depth order of the tile lists (sort stage: depth slabs + slab-wise scatter + group sort)",
                               "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "traffic": None,
                               "note": f"algorithmic bytes by SURVEY 8(d)'s accounting of a 64-bit (tile|depth) LSD radix sort: {p1s} tile-pairs x "
                                       f"24 B x {passes} digit passes, over the {sort_ms:.4f} ms CUDA-event span of the sort stage.  The stage "
                                       f"as built never makes those passes (it partitions the N Gaussians into depth slabs, scatters "
                                       f"slab-wise and sorts the small (slab, tile) groups in registers: ~16 B per pair of real traffic), "
                                       f"so it is instruction bound (64-bit sorting networks), not HBM bound"}
        del ws, img
    except Exception as e:  # noqa: BLE001
        out["stages_error"] = str(e)
    # the sorted frame as ONE CUDA graph (a viewer replays the same 13 launches with the same buffers every frame: the
    # launch-bound case graphs are for).  Captured on a side stream after a warm-up; best effort -- a capture failure is
    # reported, not fatal.
    try:
        params = capi.make_params(W, H, view.reshape(-1).tolist(), proj.reshape(-1).tolist(), (0.02, 0.02, 0.02),
                                  mode=capi.MODE_SORTED, style=capi.STYLE_NATIVE, cutoff_sigma=3.0, sh_coeffs=1,
                                  sort_depth=1, exact_bbox=1)
        mp = int(r.count_pairs(params, means, scales, opac) * 1.25) + 4096
        L = capi.lib()
        ws = torch.empty(L.b2s_workspace_bytes(n, W, H, mp) + L.b2s_state_bytes(n, W, H, mp), dtype=torch.uint8, device=device)
        img = torch.empty((H, W, 4), dtype=torch.uint8, device=device)
        kw = dict(enable_depth_sort=1, max_pairs=mp, out=img, workspace=ws)
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                r.render_rgba8(means, scales, colors, opac, view, proj, W, H, (0.02, 0.02, 0.02), **kw)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        ref_img = img.clone()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            r.render_rgba8(means, scales, colors, opac, view, proj, W, H, (0.02, 0.02, 0.02), **kw)
        img.zero_()
        graph.replay()
        torch.cuda.synchronize()
        same = bool(torch.equal(img, ref_img))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(frames):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        out["ms_per_frame_sorted_cuda_graph"] = e0.elapsed_time(e1) / frames
        out["cuda_graph_frame_identical"] = same
        del graph, ws, img
    except Exception as e:  # noqa: BLE001
        out["cuda_graph_error"] = str(e)[:300]
        torch.cuda.synchronize()
    # host-pointer path (gr::render_gaussians signature): H2D of 40 MB + render + D2H of 2 MB per frame
    hm, hs, hc, ho = (t.cpu().numpy() for t in (means, scales, colors, opac))
    bg = np.array([0.02, 0.02, 0.02], np.float32)
    r.render_gaussians(hm, hs, hc, ho, W, H, view, proj, bg, enable_depth_sort=1)
    t0 = time.perf_counter()
    for _ in range(5):
        r.render_gaussians(hm, hs, hc, ho, W, H, view, proj, bg, enable_depth_sort=1)
    out["ms_per_frame_sorted_host_buffers"] = (time.perf_counter() - t0) / 5 * 1000.0
    try:
        from oracle import cpu as ocpu
        if ocpu.have_r3ref():
            # R3: the reference's own CUDA renderer (src/renderer.cu, unmodified, sm_100a build) on the same GPU through its
            # own host-pointer API (6 blocking H2D copies + D2H per frame, :363-368,:405); viewer parameters
            # enable_depth_sort=1, depth_slices=32 (src/model_viewer_main.cpp:193-200).  Best of 10 after 2 warm-ups.
            for mode, key in ((1, "ms_per_frame_reference_cuda_r3_sliced"), (0, "ms_per_frame_reference_cuda_r3_wsum")):
                best = 1e30
                for it in range(12):
                    t0 = time.perf_counter()
                    ocpu.r3_render(hm, hs, hc, ho, view, proj, W, H, bg, depth_sort=mode, depth_slices=32)
                    dt = (time.perf_counter() - t0) * 1000.0
                    if it >= 2:
                        best = min(best, dt)
                out[key] = best
        if ocpu.have_r2ref():
            t0 = time.perf_counter()
            ocpu.r2_render(hm, hs, hc, ho, view, proj, W, H, bg, depth_sort=1)
            out["ms_per_frame_reference_cpu_1core"] = (time.perf_counter() - t0) * 1000.0
    except Exception as e:  # noqa: BLE001
        out["reference_cpu_error"] = str(e)
    out["config"] = "1M Gaussians (N,3), 960x540, enable_depth_sort=1, bg 0.02 (BASELINE configs[2])"
    return out


# ------------------------------------------------------------------------------------------
def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner at
    communicator creation), so stdout is pointed at stderr for the duration of the run and the JSON line is
    written to the saved descriptor at the end."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return saved


def ext_bench(device, n=200_000, W=960, H=540, iters=10):
    """Extension modes (csrc/splat2d.cu; absent from the reference): one view of n Gaussians, SH degree 1 (N,4,3), through
    the drop-in call with rotations= / blend=.  First correct version of these kernels (one CTA per tile, no work units):
    the numbers are here so that the gap to the tensor-core path of the reference's Gaussian model is on record."""
    r = importlib.import_module("3dgaussian_b200.renderer")
    means, scales, colors, opac = synth_gaussians(n, 4, 99, device, 0.004, 0.02)
    view, proj = synth.orbit_camera(0, 1, W, H)
    cam = r.Camera(view=torch.from_numpy(view).to(device), proj=torch.from_numpy(proj).to(device))
    quat = torch.randn((n, 4), generator=torch.Generator(device=device).manual_seed(1), device=device)
    g = torch.rand((H, W, 3), generator=torch.Generator(device=device).manual_seed(2), device=device)
    out = {"config": f"{n} Gaussians (N,4,3), {W}x{H}, one view, through render_gaussians_torch (autograd)"}
    for name, kw in (("billboard_wsum_tcgen05", {}), ("billboard_over", {"blend": "over"}),
                     ("ewa_wsum", {"rotations": quat}), ("ewa_over", {"rotations": quat, "blend": "over"})):
        leaves = [t.clone().requires_grad_(True) for t in (means, scales, colors, opac)]
        kws = dict(kw)
        if "rotations" in kws:
            kws["rotations"] = quat.clone().requires_grad_(True)
        fwd = lambda: r.render_gaussians_torch(*leaves, cam, W, H, max_gaussians=n, **kws)
        for _ in range(2):
            (fwd() * g).sum().backward()
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        with torch.no_grad():
            for _ in range(iters):
                fwd()
        e[1].record()
        for _ in range(iters):
            (fwd() * g).sum().backward()
        e[2].record()
        torch.cuda.synchronize()
        out[name] = {"fwd_ms": e[0].elapsed_time(e[1]) / iters, "fwd_bwd_ms": e[1].elapsed_time(e[2]) / iters}
    return out


def fit_scripts_bench():
    """BASELINE configs[0] and [1]: the reference's unchanged fit script
    (python/fit_multiview_stub.py:200-382, staged byte for byte under oracle/_ref/reference) three ways on this box --
      ours       : through 3dgaussian_b200/run_reference_script.py (drop-in torch_renderer / device_utils -> libb2splat);
      r1_cuda    : the reference's own torch_renderer.py with torch ops on the same GPU;
      cpu        : the reference's stock configuration (its device policy picks cpu on Linux) on the host cores, a
                   bounded number of iterations (config 2 costs tens of seconds per iteration on a CPU).
    iterations/s come from time stamps around the script's optimizer steps (start-up and image loading excluded)."""
    import subprocess
    import tempfile
    ref = os.path.join(ROOT, "oracle", "_ref", "reference")
    script = os.path.join(ref, "python", "fit_multiview_stub.py")
    if not os.path.exists(script):
        return {"unavailable": "oracle/_ref/reference not staged (make -C oracle where /root/reference exists)"}
    tex = os.path.join(ref, "assets", "scene_tex")
    c2 = os.path.join(ROOT, "tests", "golden", "c2_inputs")
    configs = {
        "config1": (["--width", "128", "--height", "128"], {"ours": 150, "r1_cuda": 150, "cpu": 12}),
        "config2": (["--width", "256", "--height", "256", "--use_sh", "--num_gaussians", "1200", "--max_gaussians", "3000",
                     "--densify_interval", "40", "--prune_interval", "40", "--masks_dir", os.path.join(c2, "masks"),
                     "--depth_dir", os.path.join(c2, "depth")], {"ours": 300, "r1_cuda": 300, "cpu": 4}),
    }
    threads = min(32, os.cpu_count() or 1)
    out = {"cpu_threads": threads}
    for name, (cfg_args, iters) in configs.items():
        row = {}
        for arm, n_it in iters.items():
            with tempfile.TemporaryDirectory() as td:
                tj = os.path.join(td, "timing.json")
                if arm == "ours":
                    cmd = [sys.executable, os.path.join(ROOT, "3dgaussian_b200", "run_reference_script.py"), "--seed", "0",
                           "--timing-json", tj, script]
                else:
                    cmd = [sys.executable, os.path.join(ROOT, "tests", "run_reference_on_torch_cuda.py"), "--seed", "0",
                           "--device", "cuda" if arm == "r1_cuda" else "cpu", "--timing-json", tj, os.path.join(ref, "python")]
                cmd += ["--targets_dir", tex, "--out_dir", os.path.join(td, "out"), "--iters", str(n_it)] + cfg_args
                env = dict(os.environ, OMP_NUM_THREADS=str(threads), MKL_NUM_THREADS=str(threads))
                t0 = time.perf_counter()
                res = subprocess.run(cmd, capture_output=True, text=True, env=env)
                wall = time.perf_counter() - t0
                if res.returncode != 0:
                    row[arm] = {"error": (res.stderr or res.stdout)[-400:]}
                    continue
                tm = json.load(open(tj))
                loss = [float(x) for x in open(os.path.join(td, "out", "loss.txt")).read().split()]
                row[arm] = {"iters": n_it, "iters_per_s": tm.get("iters_per_s"), "wall_s": wall, "loss_first": loss[0],
                            "loss_last": loss[-1]}
        if "iters_per_s" in row.get("ours", {}) and "iters_per_s" in row.get("cpu", {}) and row["cpu"]["iters_per_s"]:
            row["speedup_vs_cpu"] = row["ours"]["iters_per_s"] / row["cpu"]["iters_per_s"]
        out[name] = row
    return out


def _emit(saved_fd, line: dict):
    sys.stdout.flush()
    os.write(saved_fd, (json.dumps(line) + "\n").encode())


def main():
    args = parse_args()
    out_fd = _claim_stdout()
    if args.impl == "reference":
        run_reference(args, out_fd)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")    # NCCL's stream ahead of the compute streams it overlaps with
        torch.distributed.init_process_group("nccl", device_id=device)
    capi = importlib.import_module("3dgaussian_b200.capi")
    fit = importlib.import_module("3dgaussian_b200.fit")
    if args.render_only:
        _emit(out_fd, {"render": render_bench(device)})
        return
    if args.fit_scripts:
        _emit(out_fd, {"fit_scripts": fit_scripts_bench()})
        return
    if args.ext:
        _emit(out_fd, {"ext": ext_bench(device)})
        return
    cams = cameras(args.views, args.width, args.height)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- targets: renders of a second seeded Gaussian set (seed 4321) through our renderer ----
    gt = fit.FitDriver(args.n, args.sh, args.width, args.height, cams, device, rank=rank, world=world)
    gm, gs, gc, go = synth_gaussians(args.n, args.sh, 4321, device, args.s_lo, args.s_hi)
    gsr, gor, gcr = to_raw(gs, go, gc, args.sh)
    gt.set_params(gm, gsr, gor, gcr)
    gt.plan()
    targets, masks = {}, {}
    for i in gt.views:
        rgb, _ = gt.render_view(i)
        targets[i] = rgb.clone()
        masks[i] = (rgb.mean(dim=2) > 0.06).to(torch.float32)      # fit_multiview_stub.py:37-42
    assert not gt.check_overflow()
    del gt, gm, gs, gc, go, gsr, gor, gcr
    torch.cuda.empty_cache()

    # ---- the model being fitted (seed 1234) ----
    nv_local = len(fit.local_views(args.views, rank, world))
    view_groups = args.view_groups if args.view_groups > 0 else VIEW_GROUPS_AUTO(nv_local, args.lanes)
    drv = fit.FitDriver(args.n, args.sh, args.width, args.height, cams, device, rank=rank, world=world, lanes=args.lanes,
                        view_groups=view_groups, grad_chunks=args.grad_chunks if (args.grad_chunks > 0 and world > 1) else None,
                        comm=args.comm or None, cutoff_sigma=(args.cutoff_sigma if args.cutoff_sigma > 0 else None))
    means, scales, colors, opac = synth_gaussians(args.n, args.sh, 1234, device, args.s_lo, args.s_hi)
    sr, orr, cr = to_raw(scales, opac, colors, args.sh)
    drv.set_params(means, sr, orr, cr)
    reorder_ms = None
    if not args.no_reorder:       # setup, not part of an iteration: Gaussians stored in 3-D Morton order (locality)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        drv.reorder_spatial()
        torch.cuda.synchronize()
        reorder_ms = (time.perf_counter() - t0) * 1e3
    worst_p1 = drv.plan()
    drv.set_targets(targets, masks)
    p2 = sum(bbox_pairs(means, scales, cams[i][0], cams[i][1], args.width, args.height, 5.0) for i in drv.views)
    del means, scales, colors, opac, sr, orr, cr

    sampler = NvmlSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()          # the thread is up and polling before the warm-up; only samples inside the timed window count
    for _ in range(max(args.warmup, 3)):
        drv.step()
    if drv.check_overflow():
        drv.plan(extra_slack=1.5)
        drv.step()
        assert not drv.check_overflow(), "pair buffers overflowed twice"
    barrier()

    launches0 = capi.lib().b2s_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler:
        sampler.begin()
    if args.timeline and rank == 0:
        drv.profile = {}
    t_host0 = time.perf_counter()
    e0.record()
    n_after = []
    for it in range(args.steps):
        drv.step()
        if args.densify_every > 0 and (it + 1) % args.densify_every == 0 and it + 1 < args.steps:
            # fit_multiview_stub.py:318-325: prune / clone on the device, Adam state reset, buffers re-planned
            n_after.append(drv.densify_prune(it + 1, max_gaussians=int(args.n * 1.2), seed=1234, reorder=not args.no_reorder))
    e1.record()
    host_ms_per_step = (time.perf_counter() - t_host0) * 1e3 / args.steps     # host time to QUEUE a step (no sync inside)
    barrier()
    if sampler:
        sampler.end()
    ms = e0.elapsed_time(e1)
    timeline = None
    if drv.profile:
        prof, drv.profile = drv.profile, None
        names = ["step_begin", "preprocess_done", "views_done", "chain_done", "allreduce_done", "adam_done"]
        have = [nm for nm in names if len(prof.get(nm, [])) == args.steps]
        timeline = {}
        for a, b in zip(have[:-1], have[1:]):
            timeline[f"{a}->{b}"] = float(np.mean([x.elapsed_time(y) for x, y in zip(prof[a], prof[b])]))
        if "views_done" in prof:      # pipelined tail: when each chunk's stages finish, relative to the end of the views
            for nm in sorted(k for k in prof if k.startswith("chunk")):
                if len(prof[nm]) == args.steps:
                    timeline[f"views_done->{nm}"] = float(np.mean([x.elapsed_time(y) for x, y in zip(prof["views_done"], prof[nm])]))
        if "step_begin" in have and "adam_done" in have:
            timeline["step_begin->adam_done"] = float(np.mean([x.elapsed_time(y) for x, y in zip(prof["step_begin"], prof["adam_done"])]))
    clocks = sampler.stop() if sampler else None
    launches = capi.lib().b2s_launch_count() - launches0
    # ---- per-stage CUDA-event spans.  With one lane the stages of the timed steps themselves are bracketed; with
    # several lanes kernels of different views overlap and a span no longer times one kernel, so the SAME K steps are
    # repeated once more on a single lane with the brackets on (its total is reported as ms_per_step_one_lane).
    stages, ms_one_lane = {}, None
    if not args.no_timing and args.densify_every == 0:
        drv.active_lanes = 1
        drv.step()
        capi.timing_enable(local_rank, True)
        capi.timing_read(local_rank)
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0e.record()
        for _ in range(args.steps):
            drv.step()
        t1e.record()
        barrier()
        ms_one_lane = t0e.elapsed_time(t1e) / args.steps
        stages = capi.timing_read(local_rank)
        capi.timing_enable(local_rank, False)
        drv.active_lanes = drv.lanes
    overflowed = drv.check_overflow()
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    p2_all = torch.tensor([p2], device=device, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(p2_all)
    ms_per_step = float(t.item()) / args.steps
    value = 1000.0 / ms_per_step
    loss_last = float(drv.loss_dev.item())

    # ---- end to end: targets + masks from PINNED HOST memory every step, loss read back ----
    e2e = None
    if not args.no_e2e:
        host_t = {i: targets[i].cpu().pin_memory() for i in drv.views}
        host_m = {i: masks[i].cpu().pin_memory() for i in drv.views}
        drv.step_from_host(host_t, host_m)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            drv.step_from_host(host_t, host_m)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(dt, op=torch.distributed.ReduceOp.MAX)
        h2d = sum(host_t[i].numel() * 4 + host_m[i].numel() * 4 for i in drv.views)
        h2d_all = torch.tensor([float(h2d)], device=device, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(h2d_all)
        e2e = {"value": args.steps / float(dt.item()), "unit": "iters/s", "h2d_bytes_per_step": int(h2d_all.item()),
               "d2h_bytes_per_step": 4 * world,
               "api": "FitDriver.step_from_host: pinned-host float32 targets+masks H2D per view (double-buffered per lane), loss D2H"}
        # the same call fed with 8-bit targets / masks (decoded image bytes, converted on the device by
        # b2s_u8_to_f32): a quarter of the PCIe traffic
        host_t8 = {i: (targets[i] * 255.0).round().clamp(0, 255).to(torch.uint8).cpu().pin_memory() for i in drv.views}
        host_m8 = {i: (masks[i] * 255.0).round().to(torch.uint8).cpu().pin_memory() for i in drv.views}
        drv.step_from_host(host_t8, host_m8)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            drv.step_from_host(host_t8, host_m8)
        barrier()
        dt8 = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(dt8, op=torch.distributed.ReduceOp.MAX)
        # the same feed with the loss read back ASYNCHRONOUSLY (FitDriver.step_from_host(defer_loss=True): the copy of
        # (loss, overflow) to pinned memory is queued every step, the host consumes it one step later and never drains
        # the device); reported beside the headline, which keeps the blocking read of the reference's loop
        # (loss.item() every iteration, python/fit_multiview_stub.py:313-316)
        drv.step_from_host(host_t8, host_m8, defer_loss=True)
        drv.flush_loss()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            drv.step_from_host(host_t8, host_m8, defer_loss=True)
        drv.flush_loss()
        barrier()
        dtd = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(dtd, op=torch.distributed.ReduceOp.MAX)
        # Headline = the 8-bit feed: that is what target images ARE (the reference decodes 8-bit JPEG / PNG files,
        # python/fit_multiview_stub.py:16-34), and with several ranks pulling 2.1 GB of float32 per iteration the
        # float32 feed measures the host link, not the fit (8 GPUs: 73 vs 136 iters/s).  It stays beside it.
        f32 = {"value": e2e["value"], "unit": "iters/s", "h2d_bytes_per_step": e2e["h2d_bytes_per_step"],
               "api": e2e["api"]}
        e2e = {"value": args.steps / float(dt8.item()), "unit": "iters/s",
               "h2d_bytes_per_step": int(h2d_all.item()) // 4, "d2h_bytes_per_step": 4 * world,
               "api": "FitDriver.step_from_host: pinned-host uint8 targets+masks (the decoded image bytes) H2D per view "
                      "(double-buffered per lane), read as bytes by the fused loss kernel (b2s_fit_backward_blend_u8), loss D2H",
               "f32_targets": f32,
               "deferred_loss_read": {"value": args.steps / float(dtd.item()), "unit": "iters/s",
                                      "api": "step_from_host(defer_loss=True): same H2D feed, (loss, overflow) copied to pinned "
                                             "host memory every step without blocking, consumed one step later"}}
        del host_t8, host_m8
        del host_t, host_m

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    # ---- roofline of the dominant kernel + per-stage table (rank 0's kernels) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    sm_mhz = (clocks or {}).get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
    num_sms = int(capi.lib().b2s_sm_count())
    issue_peak = num_sms * LANES_PER_SM * sm_mhz * 1e6          # FP32 lane-instructions / s
    mufu_peak = num_sms * 16 * sm_mhz * 1e6                     # MUFU.EX2 / s
    smem_peak = num_sms * 128 * sm_mhz * 1e6 / 1e9              # GB/s of shared-memory bandwidth (128 B/clk/SM)
    nv = len(drv.views)
    p1 = float(worst_p1)                                        # worst local view (upper bound per view)
    p2_rank0 = p2
    n, sh, hw = args.n, args.sh, args.width * args.height
    n_tiles = ((args.width + 15) // 16) * ((args.height + 15) // 16)
    cs_nb = min(296, 2 * num_sms, (n + 1023) // 1024)           # counting-sort blocks (bin.cu: counting_sort_blocks)
    # Algorithmic bytes per span, DESIGN.md section 4.  A span is one view for bin / sort / blend, one STEP for the
    # batched preprocess, the chain rule and Adam.  (Round 1 modelled `sort` as a 2-pass radix sort although the timed
    # kernel is the counting-sort scatter, and `preprocess` as one view of a 64-view launch: VERDICT r1 weak #7.)
    alg_bytes = {
        # batched preprocess: parameters read once per step, 65 B (record 48 + clamp mask 1 + rect 8 + tile mask 8) per view
        "preprocess": n * (28 + 12 * sh) + n * nv * 65,
        # counting-sort stage 0 (histogram + column scan + tile scan): rect + tile mask per Gaussian, the
        # [blocks][tiles] table written once, read and rewritten once (it stays in L2)
        "bin": n * 16 + 3 * cs_nb * n_tiles * 4,
        # counting-sort stage 1 (scatter): rect + tile mask again, the table once, one 4-byte id per pair
        "sort": n * 16 + cs_nb * n_tiles * 4 + p1 * 4,
        "ranges": p1 * 8,
        "loss": hw * 48,
        "preprocess_bwd": n * (48 * nv + 2 * (28 + 12 * sh)),
        "adam": (7 + 3 * sh) * n * 28,
    }
    per_step_stages = {"preprocess", "preprocess_bwd", "adam"}
    bound_note = {"bin": "shared-memory atomics (1 per pair) + L2; HBM fraction shown for reference",
                  "sort": "shared-memory atomics + scattered 4-byte stores; HBM fraction shown for reference",
                  "preprocess": "instruction issue (~1000 instr per Gaussian*view)"}
    # SURVEY 8(d) per-unit figures of the blend: FP32 lane-instructions (+1 MUFU.EX2) per algorithmic pixel-pair,
    # each FP32 instruction counted as one FMA = 2 flop
    fp32_per_pair = {"blend_fwd": 11, "blend_bwd": 24}
    flop_per_pair = {k: 2 * v + 1 for k, v in fp32_per_pair.items()}
    # work the tensor-core formulation ISSUES per (Gaussian,tile) pair: forward one 128x32x16 tcgen05.mma per 16
    # Gaussians, backward four 128x64x16 per 128 Gaussians -> 8192 flop per pair either way; shared-memory bytes per
    # pair: operand stores by the threads + operand reads by the tensor core
    mma_flop_per_p1 = {"blend_fwd": 2 * 128 * 32 * 16 / 16, "blend_bwd": 4 * 2 * 128 * 64 * 16 / 128}
    smem_bytes_per_p1 = {"blend_fwd": 320 + 320, "blend_bwd": 64 + 128 + 64}
    tensor_peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 2250.0)))
    tensor_src = ("measured dense bf16, sustained (MEASURED_PEAKS.json)" if "bf16_tflops_sustained" in peaks
                  else "nominal 2250 TFLOP/s dense bf16 (B200_PROFILING.md fallback)")
    fp32_pipe_tflops = issue_peak * 2 / 1e12
    table = {}
    for name, (sms, spans) in stages.items():
        if spans == 0:
            continue
        per_launch_ms = sms / spans
        row = {"ms_per_step": sms / args.steps, "spans": spans, "ms_per_span": per_launch_ms}
        if name in alg_bytes:
            if name in per_step_stages:
                # the model is per STEP whatever the number of launches (the multi-GPU tail cuts the chain rule and
                # Adam into Gaussian chunks and slices: more spans, the same bytes)
                gbs = alg_bytes[name] / (row["ms_per_step"] * 1e-3) / 1e9
                row.update({"bound": "hbm", "achieved_gbs": gbs, "frac": gbs / hbm_peak, "alg_bytes_per_step": alg_bytes[name]})
            else:
                gbs = alg_bytes[name] / (per_launch_ms * 1e-3) / 1e9
                row.update({"bound": "hbm", "achieved_gbs": gbs, "frac": gbs / hbm_peak, "alg_bytes_per_span": alg_bytes[name]})
            if name in bound_note:
                row["bound_note"] = bound_note[name]
        else:
            pairs_per_span = p2_rank0 / max(nv, 1)
            sec = per_launch_ms * 1e-3
            tflops = pairs_per_span * flop_per_pair[name] / sec / 1e12
            sol_pairs = min(issue_peak / fp32_per_pair[name], mufu_peak)       # SURVEY 8(d) FP32-issue / MUFU ceiling
            row.update({"bound": "tensor", "achieved_tflops": tflops, "frac": tflops / tensor_peak,
                        "x_fp32_pipe_peak": tflops / fp32_pipe_tflops,
                        "gpairs_s": pairs_per_span / sec / 1e9,
                        "three_ways": {
                            "pairs_per_s_over_survey_fp32_sol": (pairs_per_span / sec) / sol_pairs,
                            "issued_mma_tflops": p1 * mma_flop_per_p1[name] / sec / 1e12,
                            "issued_mma_over_tensor_peak": p1 * mma_flop_per_p1[name] / sec / 1e12 / tensor_peak,
                            "smem_gbs": p1 * smem_bytes_per_p1[name] / sec / 1e9,
                            "smem_over_peak": p1 * smem_bytes_per_p1[name] / sec / 1e9 / smem_peak,
                            "note": "P2 pixel-pairs/s vs min(FP32 issue/instr per pair, MUFU) of SURVEY 8(d); MMA flop actually issued "
                                    "(8192 per (Gaussian,tile) pair, P1 = worst view) vs the measured bf16 tensor peak; shared-memory "
                                    "operand bytes (stores + tensor-core reads) vs SMs x 128 B x clock"}})
        table[name] = row
    roofline = None
    if "blend_bwd" in table:
        b = table["blend_bwd"]
        traffic = None
        try:   # dram bytes of the kernel from the committed ncu --set full capture (profiles/)
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["blend_wsum_bwd_umma_kernel"]
        except Exception:
            pass
        roofline = {"kernel": "blend_wsum_bwd_umma_kernel", "bound": "tensor",
                    "achieved": b["achieved_tflops"], "peak": tensor_peak, "unit": "TFLOP/s",
                    "frac": b["frac"], "traffic": traffic, "three_ways": b["three_ways"],
                    "note": f"algorithmic work = P2 pixel-pairs/view ({p2_rank0 / max(nv, 1):.3e}) x {flop_per_pair['blend_bwd']} flop "
                            f"(SURVEY 8d: 24 FP32 + 1 MUFU per pair) / mean span of the backward blend stage (gbuf_frag + gacc_init + "
                            f"blend_wsum_bwd_umma_kernel; the tcgen05 kernel is ~87% of it); peak = {tensor_src}. The kernel runs the "
                            f"separable sums as fp16 hi/lo tcgen05.mma 128x64x16 products (one thread per Gaussian, accumulators in "
                            f"TMEM), so the same work is {b['x_fp32_pipe_peak']:.2f}x the FP32-pipe "
                            f"peak ({fp32_pipe_tflops:.1f} TFLOP/s at {sm_mhz:.0f} MHz). `frac` follows the contract (algorithmic flop / "
                            f"tensor peak); the kernel is NOT tensor-pipe bound -- see three_ways for pairs/s vs the FP32 speed of light, "
                            f"the MMA flop actually issued and the shared-memory operand traffic (DESIGN.md section 5). "
                            f"tile-pairs P1<={p1:.3e}/view; HBM-bound stages are in roofline_stages vs {hbm_src}",
                    "share_of_step": b["ms_per_step"] / (ms_one_lane or ms_per_step),
                    "timing": "mean CUDA-event span of the kernel over the K steps repeated on one lane right after the "
                              "timed region (the timed region itself overlaps views on --lanes streams)"}

    cpu = None
    if not args.no_cpu and world == 1:
        try:
            c = cpu_fit_sample(args)
            cpu = {"value": c["iters_per_s_extrapolated"], "unit": "iters/s", "cores": c["cores"], "kind": c["kind"],
                   "sample": c["sample"], "sec_per_sample": c["sec_per_sample"], "dense_pairs_per_s": c["pairs_per_s"]}
        except Exception as e:  # noqa: BLE001
            cpu = {"error": str(e)}
    render = None
    if not args.no_render and world == 1:
        del drv, targets, masks
        torch.cuda.empty_cache()
        try:
            render = render_bench(device)
        except Exception as e:  # noqa: BLE001
            render = {"error": str(e)}

    line = {
        "metric": METRIC, "value": value, "unit": "iters/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "host_ms_per_step": host_ms_per_step, "timeline_ms": timeline,
        "comm": ("single GPU" if world == 1 else ("multimem: fused reduce-scatter + Adam + all-gather over NVLink multicast"
                                                  if drv._symm is not None else
                                                  "nccl all-reduce + replicated Adam" + (f" (multimem unavailable: {drv.comm_fallback})"
                                                                                         if getattr(drv, "comm_fallback", None) else ""))),
        "roofline": roofline, "roofline_stages": table, "ms_per_step_one_lane": ms_one_lane, "lanes": args.lanes, "view_groups": view_groups, "cpu_baseline": cpu, "render": render,
        "loss_last": loss_last, "pairs": {"P1_tile_pairs_worst_view": worst_p1, "P2_pixel_pairs_rank0_views": p2_rank0,
                                          "P2_all_ranks": float(p2_all.item())},
        "overflow": bool(overflowed), "densify": {"every": args.densify_every, "gaussians_after": n_after},
        "layout": {"gaussian_order": "generation (random)" if args.no_reorder else "3-D Morton order of the means "
                   "(FitDriver.reorder_spatial, once at setup, outside the timed region)", "reorder_ms": reorder_ms},
    }
    _emit(out_fd, line)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
