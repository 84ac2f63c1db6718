"""TEST INFRASTRUCTURE: runs the UNMODIFIED reference fit script with the reference's OWN torch_renderer.py, but on the
CUDA device (torch ops on the GPU) -- the same-box, same-seed comparison arm of tests/test_gpu_reference_script.py.

The reference's device policy returns `cpu` on Linux even with CUDA (python/device_utils.py:8-13), so a one-function
`device_utils` stand-in that returns cuda is put in front of the reference's python/ directory; every other module
(torch_renderer, the script itself) is the reference's file, byte for byte (staged under oracle/_ref/reference by
oracle/Makefile).

    python tests/run_reference_on_torch_cuda.py [--seed S] [--device cuda|cpu] <reference python dir> <script args...>
"""
import os
import runpy
import sys
import tempfile


def main():
    argv = sys.argv[1:]
    seed, device = 0, "cuda"
    while argv and argv[0] in ("--seed", "--device"):
        if argv[0] == "--seed":
            seed = int(argv[1])
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
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
