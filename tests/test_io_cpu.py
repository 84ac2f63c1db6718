"""Model / camera / target I/O (3dgaussian_b200/io.py) against the reference's formats:
gaussians npz (python/fit_multiview_stub.py:338-354, src/model_viewer_main.cpp:91-151), camera npz
(:93-111), target folders (:16-67).  CPU only."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

import scenes

io = importlib.import_module("3dgaussian_b200.io")
REF = "/root/reference/python"


def _model(n=7, sh=None, seed=0):
    r = np.random.RandomState(seed)
    m = dict(means=r.randn(n, 3).astype(np.float32), scales=r.rand(n, 3).astype(np.float32),
             colors=r.rand(n, 3).astype(np.float32), opacities=r.rand(n).astype(np.float32))
    if sh:
        m["sh_coeffs"] = r.randn(n, sh, 3).astype(np.float32)
    return m


@pytest.mark.parametrize("sh", [None, 4, 16])
def test_gaussians_npz_round_trip(tmp_path, sh):
    m = _model(sh=sh)
    p = tmp_path / "gaussians_fitted.npz"
    io.save_gaussians_npz(p, torch.from_numpy(m["means"]), m["scales"], m["colors"], m["opacities"], m.get("sh_coeffs"))
    with np.load(p) as z:      # exactly the reference's key set (fit_multiview_stub.py:338-354)
        assert sorted(z.files) == sorted(m.keys())
        assert all(z[k].dtype == np.float32 for k in z.files)
    back = io.load_gaussians_npz(p)
    for k in m:
        assert np.array_equal(back[k], m[k])


def test_gaussians_npz_validation_matches_native_loader(tmp_path):
    m = _model()
    p = tmp_path / "m.npz"
    np.savez(p, means=m["means"], scales=m["scales"], colors=m["colors"])
    with pytest.raises(ValueError, match="npz missing required arrays: means/scales/colors/opacities"):
        io.load_gaussians_npz(p)
    np.savez(p, **{**m, "means": m["means"].astype(np.float64)})
    with pytest.raises(ValueError, match="npz arrays must be float32"):
        io.load_gaussians_npz(p)
    np.savez(p, **{**m, "scales": m["scales"][:, :2]})
    with pytest.raises(ValueError, match=r"scales must be shape \(N,3\)"):
        io.load_gaussians_npz(p)
    np.savez(p, **{**m, "opacities": np.zeros((7, 2), np.float32)})
    with pytest.raises(ValueError, match=r"opacities must be shape \(N,\) or \(N,1\)"):
        io.load_gaussians_npz(p)
    np.savez(p, **{**m, "colors": m["colors"][:5]})
    with pytest.raises(ValueError, match="N mismatch"):
        io.load_gaussians_npz(p)
    np.savez(p, **{**m, "opacities": m["opacities"].reshape(-1, 1)})      # (N,1) is accepted (:125-127)
    assert io.load_gaussians_npz(p)["opacities"].shape == (7,)


def test_orbit_cameras_and_camera_npz(tmp_path):
    cams = io.orbit_cameras(5, 64, 48)
    for i, (v, p) in enumerate(cams):
        v2, p2 = scenes.orbit_camera(i, 5, 64, 48)
        assert np.array_equal(v, v2) and np.array_equal(p, p2)
    f = tmp_path / "cams.npz"
    io.save_cameras_npz(f, cams)
    back = io.load_cameras_npz(f, expected_views=5)
    assert all(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) for a, b in zip(cams, back))
    with pytest.raises(ValueError, match="camera count mismatch"):
        io.load_cameras_npz(f, expected_views=4)
    np.savez(f, view=np.zeros((1, 4, 4), np.float32))
    with pytest.raises(KeyError):
        io.load_cameras_npz(f)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_cameras_and_images_match_the_reference_module(tmp_path):
    sys.path.insert(0, REF)
    try:
        stub = importlib.import_module("fit_multiview_stub")
    finally:
        sys.path.remove(REF)
    ref = stub._make_orbit_cameras(6, 96, 64, torch.device("cpu"))
    ours = io.orbit_cameras(6, 96, 64)
    for c, (v, p) in zip(ref, ours):
        assert np.allclose(c.view.numpy(), v, atol=1e-6) and np.allclose(c.proj.numpy(), p, atol=1e-6)
    tdir = "/root/reference/assets/scene_tex"
    paths = stub._list_target_paths(__import__("pathlib").Path(tdir))
    assert [p.name for p in paths] == [p.name for p in io.list_target_paths(tdir)]
    t_ref = stub._load_image(paths[0], 80, 60)
    t_u8, masks, depths = io.load_targets_u8(tdir, 80, 60, pin=False)
    assert masks is None and depths is None
    assert np.array_equal(t_u8[0].numpy().astype(np.float32) / 255.0, t_ref)       # same decode + resize + /255
    m_ref = stub._estimate_masks([torch.from_numpy(t_ref)], 0.06)[0]
    assert torch.equal(io.estimate_masks([torch.from_numpy(t_ref)], 0.06)[0], m_ref)


def test_target_folder_with_masks(tmp_path):
    from PIL import Image
    r = np.random.RandomState(1)
    (tmp_path / "masks").mkdir()
    for name in ("01", "02"):
        Image.fromarray(r.randint(0, 255, (20, 30, 3), np.uint8), mode="RGB").save(tmp_path / f"{name}.png")
        Image.fromarray(r.randint(0, 255, (20, 30), np.uint8), mode="L").save(tmp_path / "masks" / f"{name}.png")
    t, m, d = io.load_targets_u8(tmp_path, 15, 10, masks_dir=tmp_path / "masks", depth_dir=None, pin=False)
    assert len(t) == 2 and t[0].shape == (10, 15, 3) and t[0].dtype == torch.uint8
    assert m is not None and m[1].shape == (10, 15) and d is None
    os.remove(tmp_path / "masks" / "02.png")                  # one missing file disables the set (:50-52)
    assert io.load_targets_u8(tmp_path, 15, 10, masks_dir=tmp_path / "masks", pin=False)[1] is None
    with pytest.raises(FileNotFoundError):
        io.list_target_paths(tmp_path / "masks" / "nothing")


def test_save_ppm(tmp_path):
    img = np.arange(2 * 3 * 4, dtype=np.uint8).reshape(2, 3, 4)
    io.save_ppm(tmp_path / "f.ppm", img)
    raw = (tmp_path / "f.ppm").read_bytes()
    assert raw.startswith(b"P6\n3 2\n255\n") and raw[-18:] == img[..., :3].tobytes()
