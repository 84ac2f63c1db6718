"""Seeded synthetic scenes + numpy camera helpers shared by the tests, the golden
generator and bench.py.  Camera maths restates the recipe of the reference's
fit script (/root/reference/python/fit_multiview_stub.py:70-90) and of
torch_renderer.py:24-54 in numpy float32 so the same matrices can be produced
on a box without the reference tree."""
from __future__ import annotations

import math

import numpy as np


def perspective(fovy_deg, aspect, znear, zfar):
    f = np.float32(1.0) / np.tan(np.float32(fovy_deg) * np.float32(math.pi) / np.float32(180.0) * np.float32(0.5))
    m = np.zeros((4, 4), np.float32)
    m[0, 0] = f / np.float32(aspect)
    m[1, 1] = f
    m[2, 2] = (zfar + znear) / (znear - zfar)
    m[2, 3] = (2.0 * zfar * znear) / (znear - zfar)
    m[3, 2] = -1.0
    return m


def look_at(eye, target, up):
    eye, target, up = (np.asarray(v, np.float32) for v in (eye, target, up))
    f = target - eye
    f = f / (np.linalg.norm(f) + np.float32(1e-8))
    u = up / (np.linalg.norm(up) + np.float32(1e-8))
    s = np.cross(f, u)
    s = s / (np.linalg.norm(s) + np.float32(1e-8))
    u2 = np.cross(s, f)
    m = np.eye(4, dtype=np.float32)
    m[0, :3], m[1, :3], m[2, :3] = s, u2, -f
    t = np.eye(4, dtype=np.float32)
    t[:3, 3] = -eye
    return (m @ t).astype(np.float32)


def orbit_camera(i, num_views, width, height, radius=2.5, pitch=0.2, fovy=60.0):
    yaw = (2.0 * math.pi * i) / max(1, num_views)
    eye = [radius * math.cos(pitch) * math.sin(yaw), radius * math.sin(pitch),
           radius * math.cos(pitch) * math.cos(yaw)]
    view = look_at(eye, [0, 0, 0], [0, 1, 0])
    proj = perspective(fovy, width / height, 0.01, 100.0)
    return view, proj


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
