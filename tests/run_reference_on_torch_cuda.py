"""TEST INFRASTRUCTURE: runs the UNMODIFIED reference fit script with the reference's OWN torch_renderer.py, but on the
CUDA device (torch ops on the GPU) -- the same-box, same-seed comparison arm of tests/test_gpu_reference_script.py.

The reference's device policy returns `cpu` on Linux even with CUDA (python/device_utils.py:8-13), so a one-function
`device_utils` stand-in that returns cuda is put in front of the reference's python/ directory; every other module
(torch_renderer, the script itself) is the reference's file, byte for byte (staged under oracle/_ref/reference by
oracle/Makefile).

    python tests/run_reference_on_torch_cuda.py [--seed S] [--device cuda|cpu] [--timing-json OUT] <reference python dir> <script args...>

--device cpu is the reference's stock configuration (its own device policy on Linux): bench.py --fit-scripts times it
on the GPU box's host cores as the CPU baseline of BASELINE configs[0] and [1].
"""
import os
import runpy
import sys
import tempfile


def main():
    argv = sys.argv[1:]
    seed, device, timing = 0, "cuda", None
    while argv and argv[0] in ("--seed", "--device", "--timing-json"):
        if argv[0] == "--seed":
            seed = int(argv[1])
        elif argv[0] == "--timing-json":
            timing = argv[1]
        else:
            device = argv[1]
        argv = argv[2:]
    ref_py = os.path.abspath(argv[0])
    shim = tempfile.mkdtemp()
    with open(os.path.join(shim, "device_utils.py"), "w") as f:
        f.write("import torch\n\n\ndef get_default_device():\n    return torch.device(%r)\n" % device)
    sys.path.insert(0, ref_py)
    sys.path.insert(0, shim)
    for name in ("torch_renderer", "device_utils"):
        sys.modules.pop(name, None)
    import torch
    torch.manual_seed(seed)
    script = os.path.join(ref_py, "fit_multiview_stub.py")
    sys.argv = [script] + argv[1:]
    stamps = None
    if timing:
        import importlib
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        sys.path.append(root)
        launcher = importlib.import_module("3dgaussian_b200.run_reference_script")
        stamps = launcher.install_step_timer()
    try:
        runpy.run_path(script, run_name="__main__")
    finally:
        if timing:
            launcher.write_step_timing(timing, stamps)


if __name__ == "__main__":
    main()
