"""Model / camera / target I/O and headless viewer replay around the render path.

File formats are the reference's own, so models and camera sets move between the two code bases:
  * gaussians npz   -- written by python/fit_multiview_stub.py:338-354, read by the native viewer
                       src/model_viewer_main.cpp:91-151 (means (N,3), scales (N,3), colors (N,3),
                       opacities (N,) | (N,1), optional sh_coeffs (N,K,3); float32)
  * camera npz      -- python/fit_multiview_stub.py:93-111 (view (V,4,4), proj (V,4,4))
  * target folders  -- python/fit_multiview_stub.py:16-67 (png/jpg/jpeg, bilinear resize, optional
                       masks/<stem>.png and depth/<stem>.png, masks estimated from brightness otherwise)
The viewer replay reproduces the frame loop of src/model_viewer_main.cpp:193-240 (orbit camera, fovy 60,
near 0.01, far 100, background 0.02, enable_depth_sort = 1) without a window: the model stays resident in
HBM and every frame is one b2s_render_rgba8 call (the reference re-uploads 40 MB per frame,
src/renderer.cu:363-368).
"""
from __future__ import annotations

import math
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

REQUIRED = ("means", "scales", "colors", "opacities")


# ---- gaussians npz ---------------------------------------------------------------------------
def save_gaussians_npz(path, means, scales, colors, opacities, sh_coeffs=None) -> None:
    """Same keys and dtypes as the reference's fit script (fit_multiview_stub.py:338-354): activated values,
    `colors` = the (N,3) RGB (the clamped DC band when SH is used) so the native viewer can read the file."""
    def f32(a):
        a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
        return np.ascontiguousarray(a, dtype=np.float32)
    arrays = dict(means=f32(means), scales=f32(scales), colors=f32(colors), opacities=f32(opacities))
    if sh_coeffs is not None:
        arrays["sh_coeffs"] = f32(sh_coeffs)
    np.savez(str(path), **arrays)


def load_gaussians_npz(path) -> Dict[str, np.ndarray]:
    """Validation and messages of load_gaussians_npz in src/model_viewer_main.cpp:91-151 (ValueError instead of
    the `err` string).  Returns float32 C-contiguous arrays, opacities flattened to (N,)."""
    with np.load(str(path)) as npz:
        if any(k not in npz for k in REQUIRED):
            raise ValueError("npz missing required arrays: means/scales/colors/opacities")
        a = {k: npz[k] for k in npz.files}
    for k in REQUIRED:
        if a[k].dtype.itemsize != 4 or a[k].dtype.kind != "f":
            raise ValueError("npz arrays must be float32")
    for k in ("means", "scales", "colors"):
        if a[k].ndim != 2 or a[k].shape[1] != 3:
            raise ValueError(f"{k} must be shape (N,3)")
    op = a["opacities"]
    if not (op.ndim == 1 or (op.ndim == 2 and op.shape[1] == 1)):
        raise ValueError("opacities must be shape (N,) or (N,1)")
    n = a["means"].shape[0]
    if a["scales"].shape[0] != n or a["colors"].shape[0] != n or op.shape[0] != n:
        raise ValueError("means/scales/colors/opacities N mismatch")
    out = {k: np.ascontiguousarray(a[k], dtype=np.float32) for k in ("means", "scales", "colors")}
    out["opacities"] = np.ascontiguousarray(op.reshape(n), dtype=np.float32)
    if "sh_coeffs" in a:
        out["sh_coeffs"] = np.ascontiguousarray(a["sh_coeffs"], dtype=np.float32)
    return out


# ---- cameras -----------------------------------------------------------------------------------
def perspective_np(fovy_deg: float, aspect: float, znear: float, zfar: float) -> np.ndarray:
    """torch_renderer.py:24-32 / model_viewer_main.cpp `perspective`, row-major float32."""
    f = np.float32(1.0) / np.tan(np.float32(fovy_deg) * np.float32(math.pi) / np.float32(180.0) * np.float32(0.5))
    m = np.zeros((4, 4), np.float32)
    m[0, 0] = f / np.float32(aspect)
    m[1, 1] = f
    m[2, 2] = (zfar + znear) / (znear - zfar)
    m[2, 3] = (2.0 * zfar * znear) / (znear - zfar)
    m[3, 2] = -1.0
    return m


def look_at_np(eye, target, up) -> np.ndarray:
    """torch_renderer.py:35-54 / model_viewer_main.cpp `look_at`, row-major float32, column-vector convention."""
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


def orbit_pose(yaw: float, pitch: float = 0.2, radius: float = 2.5) -> np.ndarray:
    """Eye of the viewer's orbit camera (src/model_viewer_main.cpp:230-233); defaults :188-190."""
    return np.array([radius * math.cos(pitch) * math.sin(yaw), radius * math.sin(pitch),
                     radius * math.cos(pitch) * math.cos(yaw)], np.float32)


def orbit_cameras(num_views: int, width: int, height: int, fovy: float = 60.0, pitch: float = 0.2,
                  radius: float = 2.5) -> List[Tuple[np.ndarray, np.ndarray]]:
    """_make_orbit_cameras (fit_multiview_stub.py:70-90): yaw = 2 pi i / V, near 0.01, far 100."""
    proj = perspective_np(fovy, width / height, 0.01, 100.0)
    return [(look_at_np(orbit_pose(2.0 * math.pi * i / max(1, num_views), pitch, radius), [0, 0, 0], [0, 1, 0]), proj)
            for i in range(num_views)]


def load_cameras_npz(path, expected_views: Optional[int] = None) -> List[Tuple[np.ndarray, np.ndarray]]:
    """_load_cameras (fit_multiview_stub.py:93-111), same exceptions."""
    with np.load(str(path)) as data:
        if "view" not in data or "proj" not in data:
            raise KeyError("camera npz must contain arrays: view (V,4,4), proj (V,4,4)")
        views = np.asarray(data["view"], dtype=np.float32)
        projs = np.asarray(data["proj"], dtype=np.float32)
    if expected_views is not None and (views.shape[0] != expected_views or projs.shape[0] != expected_views):
        raise ValueError("camera count mismatch with number of target images")
    return [(np.ascontiguousarray(views[i]), np.ascontiguousarray(projs[i])) for i in range(views.shape[0])]


def save_cameras_npz(path, cameras: Sequence[Tuple[np.ndarray, np.ndarray]]) -> None:
    np.savez(str(path), view=np.stack([c[0] for c in cameras]).astype(np.float32),
             proj=np.stack([c[1] for c in cameras]).astype(np.float32))


# ---- targets -----------------------------------------------------------------------------------
def list_target_paths(targets_dir) -> List[Path]:
    """_list_target_paths (fit_multiview_stub.py:26-30)."""
    d = Path(targets_dir)
    paths = sorted([*d.glob("*.png"), *d.glob("*.jpg"), *d.glob("*.jpeg")])
    if not paths:
        raise FileNotFoundError(f"No target images found in {d} (supported: png/jpg/jpeg)")
    return paths


def _decode(path: Path, width: int, height: int, mode: str) -> torch.Tensor:
    from PIL import Image
    img = Image.open(path).convert(mode).resize((width, height), Image.Resampling.BILINEAR)
    return torch.from_numpy(np.asarray(img, dtype=np.uint8).copy())


def load_targets_u8(targets_dir, width: int, height: int, masks_dir=None, depth_dir=None, pin: bool = True):
    """Decodes and resizes like _load_image / _load_gray (fit_multiview_stub.py:16-23) but keeps the 8-bit
    values: (targets [(H,W,3) uint8], masks [(H,W) uint8] | None, depths | None), pinned for H2D.  The float
    conversion (/255) happens on the device (b2s_u8_to_f32) -- FitDriver.step_from_host accepts these as is."""
    paths = list_target_paths(targets_dir)
    pin_ = (lambda t: t.pin_memory()) if (pin and torch.cuda.is_available()) else (lambda t: t)
    targets = [pin_(_decode(p, width, height, "RGB")) for p in paths]

    def optional(d):
        if d is None:
            return None
        out = []
        for p in paths:
            cand = Path(d) / f"{p.stem}.png"
            if not cand.exists():
                return None            # fit_multiview_stub.py:50-52: any missing file disables the whole set
            out.append(pin_(_decode(cand, width, height, "L")))
        return out
    return targets, optional(masks_dir), optional(depth_dir)


def estimate_masks(targets: Iterable[torch.Tensor], thresh: float = 0.06) -> List[torch.Tensor]:
    """_estimate_masks (fit_multiview_stub.py:37-42) on float targets in [0,1]."""
    return [(t.mean(dim=2) > thresh).to(torch.float32) for t in targets]


# ---- headless viewer replay ------------------------------------------------------------------
class ViewerReplay:
    """The native viewer's frame loop without a window (src/model_viewer_main.cpp:193-240): model resident on
    the device, pair buffers sized once for the worst frame of the path, one b2s_render_rgba8 per frame."""

    def __init__(self, model: Dict[str, np.ndarray], width: int = 960, height: int = 540, fovy: float = 60.0,
                 device: Optional[torch.device] = None, enable_depth_sort: int = 1,
                 background=(0.02, 0.02, 0.02), max_gaussians: int = 1_000_000):
        from . import renderer
        self._r = renderer
        self.dev = device or renderer.get_default_device()
        n = min(model["means"].shape[0], int(max_gaussians))        # `--max` of the viewer (:79)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a[:n])).to(self.dev)
        self.means, self.scales, self.colors = t(model["means"]), t(model["scales"]), t(model["colors"])
        self.opac = t(model["opacities"].reshape(-1))
        self.W, self.H, self.fovy = int(width), int(height), float(fovy)
        self.ds, self.bg = int(enable_depth_sort), tuple(float(x) for x in background)
        self.proj = perspective_np(self.fovy, self.W / self.H, 0.01, 100.0)
        self.max_pairs, self.ws = 0, None
        self.out = torch.empty((self.H, self.W, 4), dtype=torch.uint8, device=self.dev)

    def view_matrix(self, yaw: float, pitch: float = 0.2, radius: float = 2.5) -> np.ndarray:
        return look_at_np(orbit_pose(yaw, pitch, radius), [0, 0, 0], [0, 1, 0])

    def _params(self, view):
        from . import capi
        return capi.make_params(self.W, self.H, view.reshape(-1).tolist(), self.proj.reshape(-1).tolist(), self.bg,
                                mode=capi.MODE_SORTED if self.ds else capi.MODE_WSUM, style=capi.STYLE_NATIVE,
                                cutoff_sigma=3.0, sh_coeffs=1, sort_depth=1 if self.ds else 0, exact_bbox=1)

    def plan(self, views: Sequence[np.ndarray], slack: float = 1.1) -> int:
        """Sizes the pair buffers for the worst of `views` (one counting pass each; synchronises)."""
        from . import capi
        worst = max(self._r.count_pairs(self._params(v), self.means, self.scales, self.opac) for v in views)
        self.max_pairs = int(worst * slack) + 4096
        L = capi.lib()
        n = self.means.shape[0]
        need = L.b2s_workspace_bytes(n, self.W, self.H, self.max_pairs) + L.b2s_state_bytes(n, self.W, self.H, self.max_pairs)
        self.ws = torch.empty(need, dtype=torch.uint8, device=self.dev)
        return worst

    def frame(self, view: np.ndarray, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """uint8 (H,W,4) on the device; enqueued on the current stream, no synchronisation."""
        if self.ws is None:
            self.plan([view], slack=1.5)
        return self._r.render_rgba8(self.means, self.scales, self.colors, self.opac, view, self.proj, self.W, self.H,
                                    self.bg, enable_depth_sort=self.ds, max_pairs=self.max_pairs,
                                    out=self.out if out is None else out, workspace=self.ws)

    def orbit(self, frames: int, pitch: float = 0.2, radius: float = 2.5) -> List[np.ndarray]:
        """One revolution in `frames` steps; returns the frames as host uint8 arrays."""
        views = [self.view_matrix(2.0 * math.pi * i / max(1, frames), pitch, radius) for i in range(frames)]
        self.plan(views)
        outs = []
        for v in views:
            outs.append(self.frame(v).cpu().numpy().copy())
        return outs


def save_ppm(path, rgba: np.ndarray) -> None:
    """Binary PPM of an (H,W,4) or (H,W,3) uint8 frame (no image library needed on the GPU box)."""
    rgb = np.ascontiguousarray(rgba[..., :3], dtype=np.uint8)
    with open(str(path), "wb") as f:
        f.write(f"P6\n{rgb.shape[1]} {rgb.shape[0]}\n255\n".encode())
        f.write(rgb.tobytes())
