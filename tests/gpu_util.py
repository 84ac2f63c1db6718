import importlib

import numpy as np
import torch


def pkg(name=""):
    return importlib.import_module("3dgaussian_b200" + ("." + name if name else ""))


def dev():
    return torch.device("cuda", 0)


def to_dev(*arrs):
    return [torch.from_numpy(np.ascontiguousarray(a)).to(dev()) for a in arrs]


def camera(view, proj):
    r = pkg("renderer")
    v, p = to_dev(view, proj)
    return r.Camera(view=v, proj=p)
