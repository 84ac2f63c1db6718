"""Seeded synthetic test scenes.  The camera helpers live in the package (3dgaussian_b200/synth.py, shared with
bench.py) and are re-exported here for the tests and the golden generator."""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)


_synth = importlib.import_module("3dgaussian_b200.synth")
perspective, look_at, orbit_camera = _synth.perspective, _synth.look_at, _synth.orbit_camera


def make_scene(seed, n, sh=1, s_lo=0.02, s_hi=0.2, spread=0.6, edge_cases=False):
    """means U(-spread,spread)^3, log-uniform scales, opacity sigmoid(N(0,1)), colours U(0,1)
    (or SH: dc U(0,1), rest N(0,0.1))  -- SURVEY 8(d) recipe."""
    r = np.random.RandomState(seed)
    means = ((r.rand(n, 3) - 0.5) * 2 * spread).astype(np.float32)
    scales = np.exp(r.uniform(np.log(s_lo), np.log(s_hi), (n, 3))).astype(np.float32)
    opac = (1.0 / (1.0 + np.exp(-r.randn(n)))).astype(np.float32)
    if sh == 1:
        colors = r.rand(n, 3).astype(np.float32)
    else:
        colors = (0.1 * r.randn(n, sh, 3)).astype(np.float32)
        colors[:, 0, :] = r.rand(n, 3)
    if edge_cases and n >= 16:
        means[0] = [0.0, 0.0, 5.0]          # behind the camera at the orbit pose (culled: ndc.z > 1)
        means[1] = [0.0, 0.5, 2.49]         # very close to the eye
        opac[2] = -0.3                      # negative opacity -> clamp_min(0)
        opac[3] = 0.0                       # exactly zero
        scales[4] = [-0.1, -0.05, 0.1]      # negative scales -> abs()
        scales[5] = [1e-5, 1e-5, 1e-5]      # sigma clamped to 1 px
        scales[6] = [1.5, 0.01, 0.1]        # very anisotropic, huge
        means[7] = [3.0, 0.0, 0.0]          # off-screen to the side
        if sh == 1:
            colors[8] = [1.7, -0.4, 0.5]    # colour clamp
        else:
            colors[8, 0] = [1.7, -0.4, 0.5]
            colors[9, 1:] = 2.0             # SH pushes colour outside [0,1]
    return means, scales, colors, opac
