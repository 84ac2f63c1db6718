"""CPU ORACLE (test infrastructure, NOT product code) -- densify / prune.

Restates /root/reference/python/fit_multiview_stub.py:140-197 (`_densify_and_prune`) with
plain torch indexing, returning the survivor mask and the clone source indices as well, so
that the CUDA compaction (3dgaussian_b200/csrc/densify.cu) can be checked set-wise.

Parity pin: tests/golden/densify_*.npz hold outputs of the *unmodified* reference function
(tests/golden/make_golden.py); tests/test_oracle_golden.py checks this restatement
against them, including the N(0,1) jitter (same torch CPU generator, same call order).
Only tests/ may import this module.
"""
from __future__ import annotations

import torch


def densify_and_prune(means, scales_raw, op_raw, colors_raw, max_gaussians: int, densify_ratio: float,
                      prune_opacity: float):
    """Returns (means, scales_raw, op_raw, colors_raw, keep_mask, clone_src) -- clone_src indexes the
    SURVIVOR arrays (fit_multiview_stub.py:169-174)."""
    op = torch.sigmoid(op_raw)                                              # :150
    scales = torch.nn.functional.softplus(scales_raw) + 1e-3                # :151
    keep = op > prune_opacity                                               # :153
    if int(keep.sum()) < 64:                                                # :154-157
        top_keep = torch.topk(op, k=min(64, op.shape[0]), largest=True).indices
        keep = torch.zeros_like(keep, dtype=torch.bool)
        keep[top_keep] = True
    means, scales_raw, op_raw, scales = means[keep], scales_raw[keep], op_raw[keep], scales[keep]   # :159-163
    colors_raw = colors_raw[keep]
    op = torch.sigmoid(op_raw)
    n = means.shape[0]
    room = max(0, max_gaussians - n)                                        # :166
    add_n = min(room, max(0, int(n * densify_ratio)))                       # :167
    clone_src = torch.zeros(0, dtype=torch.long)
    if add_n > 0 and n > 0:                                                 # :169-174
        idx = torch.topk(op, k=min(n, add_n), largest=True).indices
        jitter = 0.25 * scales[idx] * torch.randn_like(means[idx])
        means = torch.cat([means, means[idx] + jitter], dim=0)
        scales_raw = torch.cat([scales_raw, scales_raw[idx]], dim=0)
        op_raw = torch.cat([op_raw, op_raw[idx] - 0.1], dim=0)
        colors_raw = torch.cat([colors_raw, colors_raw[idx]], dim=0)        # :182-195 (same top-k)
        clone_src = idx
    return means, scales_raw, op_raw, colors_raw, keep, clone_src
