"""Runs an UNMODIFIED reference script (e.g. python/fit_multiview_stub.py) on the B200 path.

    python 3dgaussian_b200/run_reference_script.py [--seed S] [--timing-json OUT] /path/to/fit_multiview_stub.py <script args...>

--timing-json: time stamps of the script's optimizer steps (taken by wrapping torch.optim.Adam.step from outside, the
script stays byte for byte the reference's; the device is synchronised at each stamp) -> iterations/s without the
interpreter start-up and the image loading.

Running the script directly would put its own directory (with the reference's
torch_renderer.py / device_utils.py) at sys.path[0]; this launcher executes it with runpy
after putting our drop-in `python/` directory first, so its
`from torch_renderer import ...` / `from device_utils import ...` bind to libb2splat.
"""
import os
import runpy
import sys


def install_step_timer():
    """Wraps torch.optim.Adam.step so that every optimizer step of the script leaves a (synchronised) time stamp."""
    import time

    import torch
    stamps = []
    orig = torch.optim.Adam.step

    def step(self, *a, **k):
        out = orig(self, *a, **k)
        if torch.cuda.is_available() and any(p.is_cuda for g in self.param_groups for p in g["params"]):
            torch.cuda.synchronize()
        stamps.append(time.perf_counter())
        return out

    torch.optim.Adam.step = step
    return stamps


def write_step_timing(path, stamps):
    import json
    n = len(stamps)
    out = {"optimizer_steps": n}
    if n >= 3:      # the first interval (from step 1 to step 2) already excludes start-up; skip one more for warm-up
        out["iters_per_s"] = (n - 2) / (stamps[-1] - stamps[1])
        out["seconds_per_iter"] = (stamps[-1] - stamps[1]) / (n - 2)
    with open(path, "w") as f:
        json.dump(out, f)


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    seed, timing = None, None
    while argv and argv[0] in ("--seed", "--timing-json"):
        if argv[0] == "--seed":
            seed = int(argv[1])
        else:
            timing = argv[1]
        argv = argv[2:]
    if not argv:
        raise SystemExit(__doc__)
    script = argv[0]
    here = os.path.dirname(os.path.abspath(__file__))
    shim = os.path.join(here, "python")
    for name in ("torch_renderer", "device_utils"):
        sys.modules.pop(name, None)
    sys.path.insert(0, shim)
    if seed is not None:
        import torch
        torch.manual_seed(seed)
    sys.argv = [script] + argv[1:]
    stamps = install_step_timer() if timing else None
    try:
        runpy.run_path(script, run_name="__main__")
    finally:
        if timing:
            write_step_timing(timing, stamps)


if __name__ == "__main__":
    main()
