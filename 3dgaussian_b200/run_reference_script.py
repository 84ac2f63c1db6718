"""Runs an UNMODIFIED reference script (e.g. python/fit_multiview_stub.py) on the B200 path.

    python 3dgaussian_b200/run_reference_script.py [--seed S] /path/to/fit_multiview_stub.py <script args...>

Running the script directly would put its own directory (with the reference's
torch_renderer.py / device_utils.py) at sys.path[0]; this launcher executes it with runpy
after putting our drop-in `python/` directory first, so its
`from torch_renderer import ...` / `from device_utils import ...` bind to libb2splat.
"""
import os
import runpy
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    seed = None
    if argv and argv[0] == "--seed":
        seed = int(argv[1])
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
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
