"""Seeded synthetic workloads: the orbit-camera recipe of the reference's fit script and the Gaussian sets the
benchmark and the tests render.  Product-side module: bench.py and tests/ both import it, so the
benchmark does not depend on the test tree.

Camera maths restates /root/reference/python/fit_multiview_stub.py:70-90 (`_make_orbit_cameras`) and
python/torch_renderer.py:24-54 (`perspective`, `look_at`) in numpy float32, so the same matrices can be produced on a
box without the reference tree and without a device.
"""
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
    """View i of the reference's fallback orbit (fovy 60, radius 2.5, pitch 0.2, yaw = 2 pi i / V)."""
    yaw = (2.0 * math.pi * i) / max(1, num_views)
    eye = [radius * math.cos(pitch) * math.sin(yaw), radius * math.sin(pitch),
           radius * math.cos(pitch) * math.cos(yaw)]
    view = look_at(eye, [0, 0, 0], [0, 1, 0])
    proj = perspective(fovy, width / height, 0.01, 100.0)
    return view, proj


def orbit_cameras(views, width, height):
    """[(view 16 floats, proj 16 floats)] row-major lists, the form FitDriver takes."""
    out = []
    for i in range(views):
        v, p = orbit_camera(i, views, width, height)
        out.append((v.reshape(-1).tolist(), p.reshape(-1).tolist()))
    return out


def synth_gaussians(n, sh, seed, device, s_lo=0.004, s_hi=0.02):
    """SURVEY 8(d) recipe, generated on `device` with torch: means U(-0.6,0.6)^3 (fit_multiview_stub.py:119),
    log-uniform scales, opacity sigmoid(N(0,1)), dc U(0,1), higher SH bands N(0,0.1).  Returns ACTIVATED values."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    means = (torch.rand((n, 3), generator=g, device=device) - 0.5) * 1.2
    u = torch.rand((n, 3), generator=g, device=device)
    scales = torch.exp(math.log(s_lo) + u * (math.log(s_hi) - math.log(s_lo)))
    opac = torch.sigmoid(torch.randn((n,), generator=g, device=device))
    if sh == 1:
        colors = torch.rand((n, 3), generator=g, device=device)
    else:
        colors = 0.1 * torch.randn((n, sh, 3), generator=g, device=device)
        colors[:, 0, :] = torch.rand((n, 3), generator=g, device=device)
    return means, scales, colors, opac


def to_raw(scales, opac, colors, sh):
    """Inverse activations (fit_multiview_stub.py:268-275): softplus^-1(s-1e-3), logit."""
    import torch
    s = (scales - 1e-3).clamp_min(1e-6)
    scales_raw = torch.where(s > 20.0, s, torch.log(torch.expm1(s)))
    op = opac.clamp(1e-6, 1 - 1e-6)
    op_raw = torch.log(op / (1 - op))
    if sh == 1:
        c = colors.clamp(1e-4, 1 - 1e-4)
        col_raw = torch.log(c / (1 - c))
    else:
        col_raw = colors
    return scales_raw, op_raw, col_raw
