"""CPU ORACLE (test infrastructure, NOT product code) -- weighted-sum renderer "R1".

A restatement of the reference's differentiable dense renderer
(/root/reference/python/torch_renderer.py:57-203) written as one dense
(Gaussian x pixel) evaluation with plain torch ops so that torch autograd
provides the gradient oracle.  It is dtype-parametric: run it in float32 to
mimic the reference, in float64 as the "ground truth" for the gradient tests.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product path (3dgaussian_b200/)
never does; it fails loudly when the CUDA library is missing.

Parity pin: tests/golden/r1_*.npz hold outputs of the *unmodified* reference
module generated in the build container by tests/golden/make_golden.py; the
not-gpu test-suite checks this restatement against every one of them.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch


def project(means, view, proj, width: int, height: int):
    """World -> pixel projection.  Follows torch_renderer.py:57-78 (`_project`).

    Column-vector convention (p_cam = V @ [m,1]).  Returns px, py, z_abs, valid
    and the raw camera-space z (used for the depth ordering of the sorted mode,
    src/renderer_cpu.cpp:138-146).
    """
    x, y, z = means[:, 0], means[:, 1], means[:, 2]

    def row(m, r, a, b, c, d):
        return m[r, 0] * a + m[r, 1] * b + m[r, 2] * c + m[r, 3] * d

    one = torch.ones_like(x)
    cam = [row(view, r, x, y, z, one) for r in range(4)]
    clip = [row(proj, r, cam[0], cam[1], cam[2], cam[3]) for r in range(4)]
    w = clip[3]
    w_safe = torch.where(w.abs() < 1e-8, torch.ones_like(w), w)          # :67
    ndc_x, ndc_y, ndc_z = clip[0] / w_safe, clip[1] / w_safe, clip[2] / w_safe
    px = (ndc_x * 0.5 + 0.5) * (width - 1)                               # :71
    py = (1.0 - (ndc_y * 0.5 + 0.5)) * (height - 1)                      # :72
    valid = (ndc_z >= -1.0) & (ndc_z <= 1.0) & (w != 0.0)                # :75
    z_abs = cam[2].abs().clamp_min(1e-6)                                 # :76
    return px, py, z_abs, valid, cam[2]


def camera_position(view):
    """torch_renderer.py:81-83 -- camera centre = inv(view)[:3, 3]."""
    return torch.linalg.inv(view)[:3, 3]


SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005,
         -1.0925484305920792, 0.5462742152960396)
SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658,
         0.3731763325901154, -0.4570457994644658, 1.445305721320277,
         -0.5900435899266435)


def sh_basis(dirs, n_coeffs: int):
    """Basis functions per Gaussian, (N, n_coeffs).

    Coefficients 0..3 follow the reference's non-standard "degree-1" basis
    [1, d.x, d.y, d.z] (torch_renderer.py:98-103).  Coefficients 4..15 are an
    EXTENSION (absent from the reference): the standard real-SH band-2/band-3
    polynomials in d.  They are pinned by this oracle only.
    """
    dx, dy, dz = dirs[:, 0], dirs[:, 1], dirs[:, 2]
    cols = [torch.ones_like(dx), dx, dy, dz]
    if n_coeffs > 4:
        xx, yy, zz = dx * dx, dy * dy, dz * dz
        xy, yz, xz = dx * dy, dy * dz, dx * dz
        cols += [SH_C2[0] * xy, SH_C2[1] * yz, SH_C2[2] * (2.0 * zz - xx - yy),
                 SH_C2[3] * xz, SH_C2[4] * (xx - yy)]
        if n_coeffs > 9:
            cols += [SH_C3[0] * dy * (3.0 * xx - yy), SH_C3[1] * xy * dz,
                     SH_C3[2] * dy * (4.0 * zz - xx - yy),
                     SH_C3[3] * dz * (2.0 * zz - 3.0 * xx - 3.0 * yy),
                     SH_C3[4] * dx * (4.0 * zz - xx - yy),
                     SH_C3[5] * dz * (xx - yy), SH_C3[6] * dx * (xx - 3.0 * yy)]
    return torch.stack(cols[:n_coeffs], dim=1)


def eval_colors(colors, means, view):
    """torch_renderer.py:86-106 (`_eval_colors`) + the clamp at :144."""
    if colors.ndim == 2 and colors.shape[1] == 3:
        return colors.clamp(0.0, 1.0)
    if colors.ndim == 3 and colors.shape[2] == 3 and colors.shape[1] in (4, 9, 16):
        cam = camera_position(view)
        d = cam.view(1, 3) - means
        d = d / (torch.linalg.norm(d, dim=1, keepdim=True) + 1e-8)        # :97
        basis = sh_basis(d, colors.shape[1])                             # (N,K)
        return (basis.unsqueeze(-1) * colors).sum(dim=1).clamp(0.0, 1.0)
    raise ValueError("colors must be (N,3) or SH coeffs (N,4,3)")


def screen_sigmas(scales, proj, z_abs, width: int, height: int):
    """torch_renderer.py:147-150 -- axis-aligned screen sigmas, clamped >= 1 px."""
    fx, fy = proj[0, 0].abs(), proj[1, 1].abs()
    sx = (scales[:, 0].abs() * 0.5 * width * fx / z_abs).clamp_min(1.0)
    sy = (scales[:, 1].abs() * 0.5 * height * fy / z_abs).clamp_min(1.0)
    return sx, sy


def render_r1(means, scales, colors, opacities, view, proj, width: int, height: int,
              background: Optional[torch.Tensor] = None, chunk: int = 512,
              cutoff_sigma: Optional[float] = None, tile: int = 0, pixels=None
              ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Weighted-sum blend (torch_renderer.py:109-203).  Returns (rgb, alpha, depth).

    out = clamp((bg + sum w c) / (1 + sum w)),  alpha = W/(1+W),  depth = D/(W+1e-6).
    `cutoff_sigma` (not in the reference) optionally restricts each Gaussian to
    the pixel bbox [floor(p - k s), ceil(p + k s)] -- the reference CPU/CUDA
    renderers' bbox (src/renderer_cpu.cpp:96-102) with k instead of 3 -- snapped
    outwards to `tile`-pixel tiles when tile > 0; used to restate what the
    binned CUDA path evaluates.
    `pixels` = (ix, iy) integer index tensors (P,) evaluates only those pixels of the
    width x height image (the crop oracle of the full-size parity tests: the same
    arithmetic on a subset of the reference's pixel grid) and returns (P,3), (P,), (P,).
    """
    dt, dev = means.dtype, means.device
    if background is None:
        background = torch.zeros(3, dtype=dt, device=dev)
    background = background.to(dt)
    n = means.shape[0]
    px, py, z_abs, valid, _ = project(means, view, proj, width, height)
    col = eval_colors(colors, means, view)
    sx, sy = screen_sigmas(scales, proj, z_abs, width, height)
    op = opacities.clamp_min(0.0)                                        # :175

    if pixels is None:
        ys = torch.arange(height, dtype=dt, device=dev) + 0.5            # :153-155
        xs = torch.arange(width, dtype=dt, device=dev) + 0.5
        gx = xs.view(1, 1, width)
        gy = ys.view(1, height, 1)
        ix = torch.arange(width, dtype=dt, device=dev).view(1, 1, width)
        iy = torch.arange(height, dtype=dt, device=dev).view(1, height, 1)
        hw = height * width
    else:
        ix = pixels[0].to(dt).view(1, 1, -1)
        iy = pixels[1].to(dt).view(1, 1, -1)
        gx, gy = ix + 0.5, iy + 0.5
        hw = ix.numel()

    acc_c = torch.zeros((hw, 3), dtype=dt, device=dev)
    acc_w = torch.zeros((hw,), dtype=dt, device=dev)
    acc_d = torch.zeros((hw,), dtype=dt, device=dev)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        dx = gx - px[s:e].view(-1, 1, 1)
        dy = gy - py[s:e].view(-1, 1, 1)
        ex = (dx * dx) / (sx[s:e].view(-1, 1, 1) ** 2)
        ey = (dy * dy) / (sy[s:e].view(-1, 1, 1) ** 2)
        wgt = op[s:e].view(-1, 1, 1) * torch.exp(-0.5 * (ex + ey))       # :183-184
        keep = valid[s:e].view(-1, 1, 1)
        if cutoff_sigma is not None:
            with torch.no_grad():
                x0 = torch.floor(px[s:e] - cutoff_sigma * sx[s:e]).clamp_min(0)
                x1 = torch.ceil(px[s:e] + cutoff_sigma * sx[s:e]).clamp_max(width - 1)
                y0 = torch.floor(py[s:e] - cutoff_sigma * sy[s:e]).clamp_min(0)
                y1 = torch.ceil(py[s:e] + cutoff_sigma * sy[s:e]).clamp_max(height - 1)
                if tile > 0:
                    x0 = torch.floor(x0 / tile) * tile
                    y0 = torch.floor(y0 / tile) * tile
                    x1 = torch.floor(x1 / tile) * tile + (tile - 1)
                    y1 = torch.floor(y1 / tile) * tile + (tile - 1)
                inside = ((ix >= x0.view(-1, 1, 1)) & (ix <= x1.view(-1, 1, 1)) &
                          (iy >= y0.view(-1, 1, 1)) & (iy <= y1.view(-1, 1, 1)))
            keep = keep & inside
        wgt = torch.where(keep, wgt, torch.zeros_like(wgt))              # :185
        wf = wgt.reshape(e - s, hw)
        acc_w = acc_w + wf.sum(dim=0)                                    # :188
        acc_c = acc_c + wf.t() @ col[s:e]                                # :189
        acc_d = acc_d + wf.t() @ z_abs[s:e]                              # :190

    shape = (height, width) if pixels is None else (hw,)
    wimg = acc_w.view(*shape)
    rgb = ((background.view(*([1] * len(shape)), 3) + acc_c.view(*shape, 3)) /
           (1.0 + wimg).unsqueeze(-1)).clamp(0.0, 1.0)                   # :194-196
    alpha = (wimg / (1.0 + wimg)).clamp(0.0, 1.0)                        # :201
    depth = (acc_d.view(*shape) / (wimg + 1e-6)).clamp_min(0.0)          # :202
    return rgb, alpha, depth


def render_sorted(means, scales, colors, opacities, view, proj, width: int, height: int,
                  background=None, cutoff_sigma: float = 3.0):
    """Front-to-back "over" compositing -- restatement of the reference CPU
    renderer's depth-sorted mode (src/renderer_cpu.cpp:125-217, finalize :241-257)
    as dense torch ops (float math only; the u8 quantisation is `quantise_rgba8`).

    Order: camera-space z descending, ties by index (std::stable_sort semantics;
    the reference's std::sort at :144 is unstable, so ties are "modulo order").
    Differences from R1 that this restates: z_abs = |z|+1e-6 (:172), signed
    scales (:188-189), no opacity clamp, 1/w multiply (:178-181), exact 3-sigma
    bbox (:194-200), skip a < 1e-5 (:212), a = clamp01(a) (:213).
    """
    dt, dev = means.dtype, means.device
    if background is None:
        background = torch.zeros(3, dtype=dt, device=dev)
    x, y, z = means[:, 0], means[:, 1], means[:, 2]

    def row(m, r, a, b, c, d):
        return m[r, 0] * a + m[r, 1] * b + m[r, 2] * c + m[r, 3] * d

    one = torch.ones_like(x)
    cam = [row(view, r, x, y, z, one) for r in range(4)]
    clip = [row(proj, r, cam[0], cam[1], cam[2], cam[3]) for r in range(4)]
    w = clip[3]
    inv_w = 1.0 / torch.where(w == 0, torch.ones_like(w), w)
    nx, ny, nz = clip[0] * inv_w, clip[1] * inv_w, clip[2] * inv_w
    ok = (w != 0) & (nz >= -1.0) & (nz <= 1.0)
    px = (nx * 0.5 + 0.5) * (width - 1)
    py = (1.0 - (ny * 0.5 + 0.5)) * (height - 1)
    z_abs = cam[2].abs() + 1e-6
    fx, fy = proj[0, 0].abs(), proj[1, 1].abs()
    sx = (scales[:, 0] * 0.5 * width * fx / z_abs).clamp_min(1.0)
    sy = (scales[:, 1] * 0.5 * height * fy / z_abs).clamp_min(1.0)
    x0 = torch.floor(px - cutoff_sigma * sx).clamp_min(0)
    x1 = torch.ceil(px + cutoff_sigma * sx).clamp_max(width - 1)
    y0 = torch.floor(py - cutoff_sigma * sy).clamp_min(0)
    y1 = torch.ceil(py + cutoff_sigma * sy).clamp_max(height - 1)

    order = torch.sort(cam[2], descending=True, stable=True).indices
    ix = torch.arange(width, dtype=dt, device=dev).view(1, width)
    iy = torch.arange(height, dtype=dt, device=dev).view(height, 1)
    acc = torch.zeros((height, width, 3), dtype=dt, device=dev)
    acc_a = torch.zeros((height, width), dtype=dt, device=dev)
    for i in order.tolist():
        if not bool(ok[i]):
            continue
        dx = (ix + 0.5) - px[i]
        dy = (iy + 0.5) - py[i]
        e = -0.5 * (dx * dx * (1.0 / (sx[i] * sx[i])) + dy * dy * (1.0 / (sy[i] * sy[i])))
        a = opacities[i] * torch.exp(e)
        inside = (ix >= x0[i]) & (ix <= x1[i]) & (iy >= y0[i]) & (iy <= y1[i])
        a = torch.where(inside & (a >= 1e-5), a.clamp(0.0, 1.0), torch.zeros_like(a))
        contrib = (1.0 - acc_a) * a
        contrib = torch.where(contrib > 0, contrib, torch.zeros_like(contrib))
        acc = acc + contrib.unsqueeze(-1) * colors[i].view(1, 1, 3)
        acc_a = acc_a + contrib
    a_fin = acc_a.clamp(0.0, 1.0)
    rgb = (acc + (1.0 - a_fin).unsqueeze(-1) * background.view(1, 1, 3)).clamp(0.0, 1.0)
    return rgb, a_fin


def quantise_rgba8(rgb):
    """src/renderer_cpu.cpp:252-255: u8(x*255+0.5), A = 255."""
    h, w, _ = rgb.shape
    out = torch.full((h, w, 4), 255, dtype=torch.uint8)
    out[..., :3] = (rgb.to(torch.float32) * 255.0 + 0.5).to(torch.uint8)
    return out


def adam_step(p, g, m, v, step: int, lr: float, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam defaults as used at fit_multiview_stub.py:262 (no weight
    decay, no amsgrad).  Returns updated (p, m, v); `step` is 1-based."""
    m = b1 * m + (1.0 - b1) * g
    v = b2 * v + (1.0 - b2) * g * g
    bc1 = 1.0 - b1 ** step
    bc2 = 1.0 - b2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


def fit_loss(pred, alpha, depth, tgt, mask=None, d_gt=None, silhouette_weight=0.2,
             depth_weight=0.05):
    """Per-view loss of fit_multiview_stub.py:292-303."""
    loss = torch.mean(torch.abs(pred - tgt))
    if mask is not None and silhouette_weight > 0.0:
        loss = loss + silhouette_weight * torch.mean(torch.abs(alpha - mask))
    if d_gt is not None and depth_weight > 0.0:
        d_pred = depth / (depth.max() + 1e-6)
        loss = loss + depth_weight * torch.mean(torch.abs(d_pred - d_gt))
    return loss
