"""ctypes front-ends for the compiled CPU oracles (test infrastructure, NOT product).

  libb2oracle.so  <- oracle/bins_oracle.c      (our plain-C restatement)
  libr2ref.so     <- /root/reference/src/renderer_cpu.cpp + oracle/r2_shim.cpp
                     (the unmodified reference CPU renderer; built in the build
                     container by `make -C oracle`, travels prebuilt in oracle/_ref/)
  libr3ref.so     <- /root/reference/src/renderer.cu + oracle/r3_shim.cpp, nvcc sm_100a
                     (the unmodified reference CUDA renderer: the same-box GPU baseline
                     bench.py times next to the viewer config; needs a GPU to run)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_B2O: Optional[C.CDLL] = None
_R2: Optional[C.CDLL] = None


def build(quiet: bool = True) -> None:
    """Compile bins_oracle.c, and libr2ref.so when the reference tree is present."""
    out = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def b2o() -> C.CDLL:
    global _B2O
    if _B2O is None:
        path = os.path.join(_HERE, "_build", "libb2oracle.so")
        if not os.path.exists(path):
            build()
        _B2O = C.CDLL(path)
        _B2O.b2o_project.restype = C.c_int64
        _B2O.b2o_depth_bits.restype = C.c_uint32
        _B2O.b2o_depth_bits.argtypes = [C.c_float]
    return _B2O


def r2ref() -> C.CDLL:
    global _R2
    if _R2 is None:
        path = os.path.join(_HERE, "_ref", "libr2ref.so")
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            raise FileNotFoundError("oracle/_ref/libr2ref.so missing and reference tree absent")
        _R2 = C.CDLL(path)
    return _R2


_R3: Optional[C.CDLL] = None


def have_r3ref() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libr3ref.so"))


def r3_render(means, scales, colors, opac, view, proj, width, height, bg=(0, 0, 0), depth_sort=1, depth_slices=32):
    """The reference's own gr::render_gaussians_cuda (src/renderer.cu:272-408): host pointers in, RGBA8 out,
    with its per-frame H2D copies and D2H read-back."""
    global _R3
    if _R3 is None:
        _R3 = C.CDLL(os.path.join(_HERE, "_ref", "libr3ref.so"))
    means, scales, colors, opac = _f32(means), _f32(scales), _f32(colors), _f32(opac)
    view, proj, bg = _f32(view).reshape(16), _f32(proj).reshape(16), _f32(bg)
    out = np.zeros((height, width, 4), np.uint8)
    FP = C.c_float
    rc = _R3.r3ref_render(_p(means, FP), _p(scales, FP), _p(colors, FP), _p(opac, FP), C.c_int(means.shape[0]),
                          C.c_int(width), C.c_int(height), _p(view, FP), _p(proj, FP), _p(bg, FP),
                          C.c_int(depth_sort), C.c_int(depth_slices), _p(out, C.c_uint8))
    if rc != 0:
        raise RuntimeError("reference CUDA renderer threw")
    return out


def have_r2ref() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libr2ref.so")) or \
        os.path.exists("/root/reference/src/renderer_cpu.cpp")


def bin_gaussians(means, scales, opac, view, proj, width, height, k=5.0, tile=16, style=0,
                  begin_bit=0, cull=None):
    """Full integer pipeline on the CPU.  Returns a dict of numpy arrays."""
    lib = b2o()
    means, scales, opac = _f32(means), _f32(scales), _f32(opac)
    view, proj = _f32(view).reshape(16), _f32(proj).reshape(16)
    n = means.shape[0]
    f = lambda: np.zeros(max(n, 1), np.float32)
    px, py, sx, sy, zabs, zcam = f(), f(), f(), f(), f(), f()
    bbox = np.zeros((max(n, 1), 4), np.int32)
    rect = np.zeros((max(n, 1), 4), np.int32)
    cnt = np.zeros(max(n, 1), np.int32)
    tmask = np.zeros(max(n, 1), np.uint64)
    if cull is None:
        cull = style == 0          # elliptical tile culling belongs to the torch-style weighted sum only
    FP, IP = C.c_float, C.c_int32
    total = lib.b2o_project(_p(means, FP), _p(scales, FP), _p(opac, FP), _p(view, FP), _p(proj, FP),
                            C.c_int(n), C.c_int(width), C.c_int(height), C.c_float(k), C.c_int(tile),
                            C.c_int(style), _p(px, FP), _p(py, FP), _p(sx, FP), _p(sy, FP),
                            _p(zabs, FP), _p(zcam, FP), _p(bbox, IP), _p(rect, IP), _p(cnt, IP),
                            _p(tmask, C.c_uint64), C.c_int(1 if cull else 0))
    tiles_x = (width + tile - 1) // tile
    tiles_y = (height + tile - 1) // tile
    n_tiles = tiles_x * tiles_y
    keys = np.zeros(max(total, 1), np.uint64)
    vals = np.zeros(max(total, 1), np.int32)
    lib.b2o_emit(_p(rect, IP), _p(cnt, IP), _p(tmask, C.c_uint64), _p(zcam, FP), C.c_int(n), C.c_int(tiles_x),
                 _p(keys, C.c_uint64), _p(vals, IP))
    keys_unsorted, vals_unsorted = keys[:total].copy(), vals[:total].copy()
    tile_bits = max(1, int(np.ceil(np.log2(max(n_tiles, 2)))))
    lib.b2o_sort(_p(keys, C.c_uint64), _p(vals, IP), C.c_int64(total), C.c_int(begin_bit),
                 C.c_int(32 + tile_bits))
    ranges = np.zeros((n_tiles, 2), np.int32)
    lib.b2o_ranges(_p(keys, C.c_uint64), C.c_int64(total), C.c_int(n_tiles), _p(ranges, IP))
    return dict(px=px[:n], py=py[:n], sx=sx[:n], sy=sy[:n], zabs=zabs[:n], zcam=zcam[:n],
                bbox=bbox[:n], rect=rect[:n], cnt=cnt[:n], total=int(total),
                keys_unsorted=keys_unsorted, vals_unsorted=vals_unsorted,
                keys=keys[:total], vals=vals[:total], ranges=ranges,
                tiles_x=tiles_x, tiles_y=tiles_y)


def blend_wsum(bins, op, col, width, height, bg, tile=16):
    lib = b2o()
    FP, IP = C.c_float, C.c_int32
    op, col, bg = _f32(op), _f32(col), _f32(bg)
    rgb = np.zeros((height, width, 3), np.float32)
    alpha = np.zeros((height, width), np.float32)
    depth = np.zeros((height, width), np.float32)
    vals = np.ascontiguousarray(bins["vals"] if bins["total"] else np.zeros(1, np.int32))
    ranges = np.ascontiguousarray(bins["ranges"])
    lib.b2o_blend_wsum(_p(bins["px"], FP), _p(bins["py"], FP), _p(bins["sx"], FP), _p(bins["sy"], FP),
                       _p(bins["zabs"], FP), _p(op, FP), _p(col, FP), _p(vals, IP), _p(ranges, IP),
                       C.c_int(width), C.c_int(height), C.c_int(tile), _p(bg, FP),
                       _p(rgb, FP), _p(alpha, FP), _p(depth, FP))
    return rgb, alpha, depth


def r2_render(means, scales, colors, opac, view, proj, width, height, bg=(0, 0, 0), depth_sort=1):
    """The reference's own gr::render_gaussians_cpu (src/renderer_cpu.cpp:34-260)."""
    lib = r2ref()
    means, scales, colors, opac = _f32(means), _f32(scales), _f32(colors), _f32(opac)
    view, proj, bg = _f32(view).reshape(16), _f32(proj).reshape(16), _f32(bg)
    out = np.zeros((height, width, 4), np.uint8)
    FP = C.c_float
    rc = lib.r2ref_render(_p(means, FP), _p(scales, FP), _p(colors, FP), _p(opac, FP),
                          C.c_int(means.shape[0]), C.c_int(width), C.c_int(height), _p(view, FP),
                          _p(proj, FP), _p(bg, FP), C.c_int(depth_sort), _p(out, C.c_uint8))
    if rc != 0:
        raise RuntimeError("reference renderer threw")
    return out
