"""ctypes binding of libb2splat.so (include/b2splat.h).

There is NO fallback: if the library is missing or fails to load, importing the
render path raises.  Build it with `python 3dgaussian_b200/build.py`.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2S_LIB_PATH") or os.path.join(_HERE, "csrc", "libb2splat.so")   # override: A/B experiments

MODE_WSUM, MODE_SORTED = 0, 1
STYLE_TORCH, STYLE_NATIVE = 0, 1
BLEND_WSUM, BLEND_OVER = 0, 1      # b2s_forward_ext / b2s_backward_ext (extension modes)
ACT_SCALES_SOFTPLUS, ACT_OPACITY_SIGMOID, ACT_COLORS_SIGMOID = 1, 2, 4
TILE = 16

EXPORTS = [
    "b2s_create", "b2s_destroy", "b2s_last_error", "b2s_version", "b2s_state_bytes", "b2s_workspace_bytes",
    "b2s_count_pairs", "b2s_forward", "b2s_backward", "b2s_state_info", "b2s_render_rgba8",
    "b2s_render_rgba8_host", "b2s_dump_bins", "b2s_sort_tmp_bytes", "b2s_sort_pairs", "b2s_fit_loss",
    "b2s_adam_step", "b2s_view_block_bytes", "b2s_pack_views", "b2s_backward_blend", "b2s_fit_backward_blend", "b2s_backward_params",
    "b2s_prepared_view_bytes", "b2s_preprocess_views", "b2s_forward_prepared", "b2s_u8_to_f32",
    "b2s_densify_workspace_bytes", "b2s_densify_prune", "b2s_launch_count", "b2s_num_stages", "b2s_stage_name", "b2s_timing_enable", "b2s_timing_read",
    "b2s_last_ticket", "b2s_ticket_info", "b2s_path_counts", "b2s_sm_count", "b2s_adam_step_guarded", "b2s_backward_params_range",
    "b2s_forward_ext", "b2s_backward_ext", "b2s_adam_step_multimem", "b2s_reduce_tail_multimem", "b2s_fit_backward_blend_u8", "b2s_multimem_share",
]


class Params(C.Structure):
    """struct b2s_params -- mirrors gr::RenderParams (reference include/gr/gaussian_types.h:24-46)."""
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32),
        ("view", C.c_float * 16), ("proj", C.c_float * 16), ("background", C.c_float * 3),
        ("enable_depth_sort", C.c_int32), ("depth_slices", C.c_int32), ("force_cpu", C.c_int32),
        ("style", C.c_int32), ("cutoff_sigma", C.c_float), ("sh_coeffs", C.c_int32),
        ("sort_depth", C.c_int32), ("act_flags", C.c_int32), ("exact_bbox", C.c_int32),
        ("background_dev", C.c_void_p), ("keep_depth", C.c_int32), ("reserved_", C.c_int32),
    ]


class B2SError(RuntimeError):
    pass


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B2SError(f"{LIB_PATH} not built (run `python 3dgaussian_b200/build.py`); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp, i32, i64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t
        PP = C.POINTER(Params)
        L.b2s_create.restype = vp
        L.b2s_create.argtypes = [i32]
        L.b2s_destroy.restype = None
        L.b2s_destroy.argtypes = [vp]
        L.b2s_last_error.restype = C.c_char_p
        L.b2s_version.restype = C.c_char_p
        L.b2s_state_bytes.restype = sz
        L.b2s_state_bytes.argtypes = [i32, i32, i32, i64]
        L.b2s_workspace_bytes.restype = sz
        L.b2s_workspace_bytes.argtypes = [i32, i32, i32, i64]
        L.b2s_count_pairs.restype = i32
        L.b2s_count_pairs.argtypes = [vp, PP, vp, vp, vp, i32, C.POINTER(i64), vp, sz, vp]
        L.b2s_forward.restype = i32
        L.b2s_forward.argtypes = [vp, PP, vp, vp, vp, vp, i32, i64, vp, vp, vp, vp, sz, vp, sz, vp]
        L.b2s_backward.restype = i32
        L.b2s_backward.argtypes = [vp, PP, vp, vp, vp, vp, i32, i64, vp, vp, vp, vp, vp, sz, vp, vp, vp, vp, i32, vp]
        L.b2s_forward_ext.restype = i32
        L.b2s_forward_ext.argtypes = [vp, PP, vp, vp, vp, vp, vp, i32, i64, i32, C.c_float, vp, vp, vp, vp, sz, vp, sz, vp]
        L.b2s_backward_ext.restype = i32
        L.b2s_backward_ext.argtypes = [vp, PP, vp, vp, vp, vp, vp, i32, i64, i32, C.c_float, vp, vp, vp, vp, vp, sz,
                                       vp, vp, vp, vp, vp, vp]
        L.b2s_view_block_bytes.restype = sz
        L.b2s_view_block_bytes.argtypes = []
        L.b2s_pack_views.restype = i32
        L.b2s_pack_views.argtypes = [PP, i32, vp]
        L.b2s_backward_blend.restype = i32
        L.b2s_backward_blend.argtypes = [vp, PP, i32, i64, vp, vp, vp, vp, vp, sz, vp, vp]
        L.b2s_fit_backward_blend.restype = i32
        L.b2s_fit_backward_blend.argtypes = [vp, PP, i32, i64, vp, vp, vp, C.c_float, C.c_float, C.c_float, vp, vp, vp, vp, sz, vp, vp]
        L.b2s_fit_backward_blend_u8.restype = i32
        L.b2s_fit_backward_blend_u8.argtypes = L.b2s_fit_backward_blend.argtypes
        L.b2s_prepared_view_bytes.restype = sz
        L.b2s_prepared_view_bytes.argtypes = [i32]
        L.b2s_preprocess_views.restype = i32
        L.b2s_preprocess_views.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, i32, vp, vp]
        L.b2s_forward_prepared.restype = i32
        L.b2s_forward_prepared.argtypes = [vp, PP, vp, i32, i64, vp, vp, vp, vp, sz, vp, sz, vp]
        L.b2s_backward_params.restype = i32
        L.b2s_backward_params.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, i32, vp]
        L.b2s_backward_params_range.restype = i32
        L.b2s_backward_params_range.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, i32, vp]
        L.b2s_state_info.restype = i32
        L.b2s_state_info.argtypes = [vp, vp, i32, i32, i32, i64, C.POINTER(i64), vp]
        L.b2s_render_rgba8.restype = i32
        L.b2s_render_rgba8.argtypes = [vp, PP, vp, vp, vp, vp, i32, i64, vp, vp, sz, vp]
        L.b2s_render_rgba8_host.restype = i32
        L.b2s_render_rgba8_host.argtypes = [vp, PP, vp, vp, vp, vp, i32, vp]
        L.b2s_dump_bins.restype = i32
        L.b2s_dump_bins.argtypes = [vp, PP, vp, vp, vp, i32, i64] + [vp] * 13 + [vp, sz, vp]
        L.b2s_sort_tmp_bytes.restype = sz
        L.b2s_sort_tmp_bytes.argtypes = [i64]
        L.b2s_sort_pairs.restype = i32
        L.b2s_sort_pairs.argtypes = [vp, vp, vp, vp, vp, i64, i32, i32, vp, sz, vp]
        L.b2s_fit_loss.restype = i32
        L.b2s_fit_loss.argtypes = [vp, vp, vp, vp, vp, i32, i32, C.c_float, C.c_float, vp, vp, vp, vp]
        L.b2s_u8_to_f32.restype = i32
        L.b2s_u8_to_f32.argtypes = [vp, vp, vp, i64, vp]
        L.b2s_adam_step.restype = i32
        L.b2s_adam_step.argtypes = [vp, vp, vp, vp, vp, i64, i32, C.c_float, C.c_float, C.c_float, C.c_float,
                                    i64, i64, C.c_float, i64, i64, C.c_float, vp]
        L.b2s_adam_step_guarded.restype = i32
        L.b2s_adam_step_guarded.argtypes = [vp, vp, vp, vp, vp, i64, i32, C.c_float, C.c_float, C.c_float, C.c_float,
                                            i64, i64, C.c_float, i64, i64, C.c_float, vp, vp, vp]
        L.b2s_adam_step_multimem.restype = i32
        L.b2s_adam_step_multimem.argtypes = [vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, C.c_float, C.c_float, C.c_float,
                                             C.c_float, i64, i64, C.c_float, i64, i64, C.c_float, vp, vp, vp]
        L.b2s_multimem_share.restype = i32
        L.b2s_multimem_share.argtypes = [i64, i32, i32, C.POINTER(i64), C.POINTER(i64)]
        L.b2s_reduce_tail_multimem.restype = i32
        L.b2s_reduce_tail_multimem.argtypes = [vp, vp, vp, i32, vp]
        L.b2s_densify_workspace_bytes.restype = sz
        L.b2s_densify_workspace_bytes.argtypes = [i32]
        L.b2s_densify_prune.restype = i32
        L.b2s_densify_prune.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, C.c_double, C.c_float, C.c_uint64, C.c_uint64,
                                        vp, vp, vp, vp, C.POINTER(C.c_int), vp, sz, vp]
        L.b2s_launch_count.restype = i64
        L.b2s_launch_count.argtypes = []
        L.b2s_last_ticket.restype = i64
        L.b2s_last_ticket.argtypes = [vp]
        L.b2s_ticket_info.restype = i32
        L.b2s_ticket_info.argtypes = [vp, i64, C.POINTER(i64)]
        L.b2s_path_counts.restype = None
        L.b2s_path_counts.argtypes = [C.POINTER(i64)]
        L.b2s_sm_count.restype = i32
        L.b2s_sm_count.argtypes = []
        L.b2s_num_stages.restype = i32
        L.b2s_stage_name.restype = C.c_char_p
        L.b2s_stage_name.argtypes = [i32]
        L.b2s_timing_enable.restype = i32
        L.b2s_timing_enable.argtypes = [vp, i32]
        L.b2s_timing_read.restype = i32
        L.b2s_timing_read.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(i64)]
        _lib = L
    return _lib


def timing_enable(device_index: int, on: bool) -> None:
    check(lib().b2s_timing_enable(ctx(device_index), 1 if on else 0))


def timing_read(device_index: int) -> dict:
    """{stage name: (milliseconds, spans)} accumulated since the last read."""
    ns = lib().b2s_num_stages()
    ms = (C.c_float * ns)()
    cnt = (C.c_int64 * ns)()
    check(lib().b2s_timing_read(ctx(device_index), ms, cnt))
    return {lib().b2s_stage_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(ns)}


def multimem_share(count: int, rank: int, world: int) -> tuple:
    """[lo, hi) of a slice of `count` floats owned by `rank` in the fused multi-GPU Adam step (host arithmetic only)."""
    lo, hi = C.c_int64(0), C.c_int64(0)
    check(lib().b2s_multimem_share(int(count), int(rank), int(world), C.byref(lo), C.byref(hi)))
    return int(lo.value), int(hi.value)


def path_counts() -> dict:
    """Launches so far of the weighted-sum blend kernel families (tests prove with it which path ran)."""
    out = (C.c_int64 * 4)()
    lib().b2s_path_counts(out)
    return {"fwd_tcgen05": int(out[0]), "fwd_other": int(out[1]), "bwd_tcgen05": int(out[2]), "bwd_other": int(out[3])}


def ticket_info(device_index: int, ticket: Optional[int] = None) -> tuple:
    """(pairs needed, pairs kept, overflow flag) of a forward call; waits for that call's binning kernels only."""
    c = ctx(device_index)
    if ticket is None:
        ticket = lib().b2s_last_ticket(c)
    info = (C.c_int64 * 3)()
    check(lib().b2s_ticket_info(c, ticket, info))
    return int(info[0]), int(info[1]), int(info[2])


def check(rc: int) -> None:
    if rc != 0:
        raise B2SError(f"libb2splat error {rc}: {lib().b2s_last_error().decode()}")


_ctx = {}


def ctx(device_index: int):
    """One b2s_ctx per device per process."""
    if device_index not in _ctx:
        h = lib().b2s_create(device_index)
        if not h:
            raise B2SError(lib().b2s_last_error().decode())
        _ctx[device_index] = h
    return _ctx[device_index]


def make_params(width, height, view, proj, background=(0.0, 0.0, 0.0), mode=MODE_WSUM, style=STYLE_TORCH,
                cutoff_sigma=5.0, sh_coeffs=1, sort_depth=0, act_flags=0, exact_bbox=0, background_dev=None,
                keep_depth=0) -> Params:
    """view/proj: 16 floats row-major (any iterable).  background_dev: optional device address of 3 floats that
    overrides `background` inside the kernels (the caller keeps that memory alive while the work is queued)."""
    p = Params()
    p.width, p.height = int(width), int(height)
    p.view[:] = [float(x) for x in view]
    p.proj[:] = [float(x) for x in proj]
    p.background[:] = [float(x) for x in background]
    p.enable_depth_sort = int(mode)
    p.depth_slices = 32
    p.force_cpu = 0
    p.style = int(style)
    p.cutoff_sigma = float(cutoff_sigma)
    p.sh_coeffs = int(sh_coeffs)
    p.sort_depth = int(sort_depth)
    p.act_flags = int(act_flags)
    p.exact_bbox = int(exact_bbox)
    p.background_dev = background_dev
    p.keep_depth = int(keep_depth)
    p.reserved_ = 0
    return p
