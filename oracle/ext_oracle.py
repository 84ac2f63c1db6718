"""CPU ORACLE (test infrastructure, NOT product code) -- extension modes named by `north_star`
that the reference does not have: `rotations` + EWA 2-D covariance, and a DIFFERENTIABLE
front-to-back ("over") compositing.

PARITY UNPINNED against the reference by construction (no rotation, no
covariance projection and no differentiable sorted mode exist anywhere in /root/reference).
What IS pinned:
  * rotations=None, blend="wsum" is exactly oracle/r1_oracle.render_r1 (the pinned R1 port);
  * rotations=None, blend="over" follows the compositing rule of the reference CPU renderer's
    depth-sorted mode (src/renderer_cpu.cpp:125-217, restated and pinned as
    r1_oracle.render_sorted): z-descending order, exact k-sigma pixel bbox, a < 1e-5 skipped,
    contrib = T a.  tests/test_ext_oracle_cpu.py checks the two against each other.
Everything is dense torch (Gaussian x pixel), dtype-parametric, so autograd in float64 is the
gradient truth for the CUDA extension kernels (csrc/splat2d.cu).

Only tests/ may import this module.

Definitions (extension semantics, chosen here):
  EWA:   t = V[m,1] (camera space), clip = P[t,1], px/py as torch_renderer.py:57-78.
         J = d(px,py)/dt (exact Jacobian of that projection for any proj matrix),
         M = R(q/|q|) diag(|s|),  T = J V[:3,:3] M,  cov2 = T T^T + dilation*I  (dilation 0.3 px^2),
         conic = cov2^-1,  w = op exp(-0.5 (A dx^2 + 2 B dx dy + C dy^2)),
         bbox half-widths k sqrt(cov2_xx), k sqrt(cov2_yy).
  over:  a_i = op_i G_i inside the Gaussian's exact pixel bbox, 0 if < 1e-5, capped at ALPHA_MAX
         (the reference clamps at 1.0; the cap keeps the transmittance a non-zero product so
         that the backward pass can divide by 1 - a: at most 1e-6 on the image);
         T_i = prod_{k<i} (1 - a_k);  C = sum T_i a_i c_i;  rgb = clamp(C + T_end bg);
         alpha = clamp(1 - T_end);  depth = sum T_i a_i z_i  (expected depth, not normalised).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import r1_oracle as r1

ALPHA_MAX = 0.999999
EWA_DILATION = 0.3


def quat_to_rot(q):
    """(N,4) quaternions (w,x,y,z), normalised here -> (N,3,3)."""
    q = q / (q.norm(dim=1, keepdim=True) + 1e-12)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    rows = [
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], dim=1),
        torch.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)], dim=1),
        torch.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], dim=1),
    ]
    return torch.stack(rows, dim=1)


def splats(means, scales, rotations, view, proj, width: int, height: int, dilation: float = EWA_DILATION):
    """Per-Gaussian 2-D splats: px, py, z_abs, valid, cam_z, conic (A,B,C) and bbox sigmas (sx, sy)."""
    px, py, z_abs, valid, camz = r1.project(means, view, proj, width, height)
    if rotations is None:
        sx, sy = r1.screen_sigmas(scales, proj, z_abs, width, height)
        return px, py, z_abs, valid, camz, (1.0 / (sx * sx), torch.zeros_like(sx), 1.0 / (sy * sy)), (sx, sy)
    n = means.shape[0]
    ones = torch.ones((n, 1), dtype=means.dtype, device=means.device)
    t = (torch.cat([means, ones], dim=1) @ view.t())[:, :3]                  # camera space
    clip = torch.cat([t, ones], dim=1) @ proj.t()
    w = clip[:, 3]
    w_safe = torch.where(w.abs() < 1e-8, torch.ones_like(w), w)
    a, b = 0.5 * (width - 1), -0.5 * (height - 1)
    p3 = proj[:, :3]
    j0 = a * (p3[0].view(1, 3) * w.view(-1, 1) - clip[:, 0:1] * p3[3].view(1, 3)) / (w_safe * w_safe).view(-1, 1)
    j1 = b * (p3[1].view(1, 3) * w.view(-1, 1) - clip[:, 1:2] * p3[3].view(1, 3)) / (w_safe * w_safe).view(-1, 1)
    J = torch.stack([j0, j1], dim=1)                                          # (N,2,3)
    M = quat_to_rot(rotations) * scales.abs().view(n, 1, 3)                   # R diag(s)
    T = J @ view[:3, :3].view(1, 3, 3) @ M                                    # (N,2,3)
    cov = T @ T.transpose(1, 2)
    c11, c12, c22 = cov[:, 0, 0] + dilation, cov[:, 0, 1], cov[:, 1, 1] + dilation
    det = c11 * c22 - c12 * c12
    return px, py, z_abs, valid, camz, (c22 / det, -c12 / det, c11 / det), (c11.sqrt(), c22.sqrt())


def render_ext(means, scales, rotations, colors, opacities, view, proj, width: int, height: int,
               background: Optional[torch.Tensor] = None, blend: str = "wsum",
               cutoff_sigma: Optional[float] = None, dilation: float = EWA_DILATION):
    """Returns (rgb, alpha, depth).  blend = "wsum" (torch_renderer.py:183-203 semantics) or "over"."""
    dt, dev = means.dtype, means.device
    if background is None:
        background = torch.zeros(3, dtype=dt, device=dev)
    background = background.to(dt)
    px, py, z_abs, valid, camz, (A, B, C), (sx, sy) = splats(means, scales, rotations, view, proj, width, height, dilation)
    col = r1.eval_colors(colors, means, view)
    ix = torch.arange(width, dtype=dt, device=dev).view(1, 1, width)
    iy = torch.arange(height, dtype=dt, device=dev).view(1, height, 1)
    dx = (ix + 0.5) - px.view(-1, 1, 1)
    dy = (iy + 0.5) - py.view(-1, 1, 1)
    power = -0.5 * (A.view(-1, 1, 1) * dx * dx + 2.0 * B.view(-1, 1, 1) * dx * dy + C.view(-1, 1, 1) * dy * dy)
    g = torch.exp(power)
    keep = valid.view(-1, 1, 1)
    if cutoff_sigma is not None:
        with torch.no_grad():
            x0 = torch.floor(px - cutoff_sigma * sx).clamp_min(0)
            x1 = torch.ceil(px + cutoff_sigma * sx).clamp_max(width - 1)
            y0 = torch.floor(py - cutoff_sigma * sy).clamp_min(0)
            y1 = torch.ceil(py + cutoff_sigma * sy).clamp_max(height - 1)
            inside = ((ix >= x0.view(-1, 1, 1)) & (ix <= x1.view(-1, 1, 1)) &
                      (iy >= y0.view(-1, 1, 1)) & (iy <= y1.view(-1, 1, 1)))
        keep = keep & inside
    if blend == "wsum":
        w = opacities.clamp_min(0.0).view(-1, 1, 1) * g
        w = torch.where(keep, w, torch.zeros_like(w))
        W = w.sum(dim=0)
        Cc = torch.einsum("nhw,nc->hwc", w, col)
        D = torch.einsum("nhw,n->hw", w, z_abs)
        rgb = ((background.view(1, 1, 3) + Cc) / (1.0 + W).unsqueeze(-1)).clamp(0.0, 1.0)
        return rgb, (W / (1.0 + W)).clamp(0.0, 1.0), (D / (W + 1e-6)).clamp_min(0.0)
    if blend != "over":
        raise ValueError("blend must be 'wsum' or 'over'")
    a = opacities.view(-1, 1, 1) * g
    a = torch.where(keep & (a >= 1e-5), a.clamp(0.0, ALPHA_MAX), torch.zeros_like(a))
    order = torch.sort(camz, descending=True, stable=True).indices          # nearest first (camera looks down -z)
    a = a[order]
    one_minus = 1.0 - a
    T_excl = torch.cumprod(torch.cat([torch.ones_like(a[:1]), one_minus[:-1]], dim=0), dim=0)
    contrib = T_excl * a
    T_end = T_excl[-1] * one_minus[-1]
    Cc = torch.einsum("nhw,nc->hwc", contrib, col[order])
    D = torch.einsum("nhw,n->hw", contrib, z_abs[order])
    rgb = (Cc + T_end.unsqueeze(-1) * background.view(1, 1, 3)).clamp(0.0, 1.0)
    return rgb, (1.0 - T_end).clamp(0.0, 1.0), D
