"""GPU tests of the fit path: FitDriver (fused activations + loss + backward + Adam) and the
drop-in autograd entry against the CPU oracle running the reference's own loop semantics
(python/fit_multiview_stub.py:265-311)."""
import importlib
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

import scenes
from conftest import rel_l2
from gpu_util import dev, pkg
from oracle import r1_oracle as r1

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _setup(sh, n=203, V=3, W=48, H=32, seed=5, s_lo=0.03, s_hi=0.15):
    means, scales, colors, opac = scenes.make_scene(seed, n, sh=sh, s_lo=s_lo, s_hi=s_hi)
    rng = np.random.RandomState(seed)
    scales_raw = np.log(np.expm1(np.maximum(scales - 1e-3, 1e-4))).astype(np.float32)
    op_raw = np.log(opac / (1 - opac)).astype(np.float32)
    col_raw = colors if sh > 1 else np.log(np.clip(colors, 1e-3, 1 - 1e-3) / (1 - np.clip(colors, 1e-3, 1 - 1e-3))).astype(np.float32)
    cams = [scenes.orbit_camera(i, V, W, H) for i in range(V)]
    tgts = [rng.rand(H, W, 3).astype(np.float32) for _ in range(V)]
    masks = [(rng.rand(H, W) > 0.5).astype(np.float32) for _ in range(V)]
    return dict(n=n, sh=sh, V=V, W=W, H=H, means=means, scales_raw=scales_raw, op_raw=op_raw, col_raw=col_raw,
                cams=cams, tgts=tgts, masks=masks)


def _oracle_loss(S, leaves, dt):
    t = lambda a: torch.from_numpy(a).to(dt)
    m, sr, orr, cr = leaves
    sc = torch.nn.functional.softplus(sr) + 1e-3
    op = torch.sigmoid(orr)
    col = torch.sigmoid(cr) if S["sh"] == 1 else cr
    total = torch.zeros((), dtype=dt)
    for i in range(S["V"]):
        rgb, alpha, depth = r1.render_r1(m, sc, col, op, t(S["cams"][i][0]), t(S["cams"][i][1]), S["W"], S["H"])
        total = total + r1.fit_loss(rgb, alpha, depth, t(S["tgts"][i]), t(S["masks"][i]), None, silhouette_weight=0.2)
    data = total / S["V"]
    return data, data + 1e-3 * op.mean() + 1e-3 * sc.mean()


def _driver(S, **kw):
    fit = pkg("fit")
    cams = [(v.reshape(-1).tolist(), p.reshape(-1).tolist()) for v, p in S["cams"]]
    d = fit.FitDriver(S["n"], S["sh"], S["W"], S["H"], cams, dev(), **kw)
    t = lambda a: torch.from_numpy(a).to(dev())
    d.set_params(t(S["means"]), t(S["scales_raw"]), t(S["op_raw"]), t(S["col_raw"]))
    d.plan()
    views = d.views
    d.set_targets({i: t(S["tgts"][i]) for i in views}, {i: t(S["masks"][i]) for i in views})
    return d


@pytest.mark.parametrize("sh", [1, 4, 16])
def test_fit_step_gradients_and_adam_match_oracle(sh):
    S = _setup(sh)
    d = _driver(S)
    p0 = d.p.clone()
    loss_dev = d.step()
    assert not d.check_overflow()
    dt = torch.float64
    leaves = [torch.from_numpy(S[k]).to(dt).requires_grad_(True) for k in ("means", "scales_raw", "op_raw", "col_raw")]
    data, full = _oracle_loss(S, leaves, dt)
    data.backward()
    ref_g = torch.cat([l.grad.reshape(-1) for l in leaves]).numpy()
    assert abs(float(loss_dev.item()) - float(data)) <= 1e-5
    n = S["n"]
    got = [v.cpu().numpy() for v in d.grad_views()]
    for name, gv, (a, b) in zip(("means", "scales", "opac", "colors"), got,
                                ((0, 3 * n), (3 * n, 6 * n), (6 * n, 7 * n), (7 * n, None))):
        assert rel_l2(gv, ref_g[a:b]) <= 1e-3, name
    # Adam with the regulariser gradients == torch Adam on the full loss
    leaves2 = [torch.nn.Parameter(torch.from_numpy(S[k]).to(dt)) for k in ("means", "scales_raw", "op_raw", "col_raw")]
    opt = torch.optim.Adam(leaves2, lr=0.02)
    _, full2 = _oracle_loss(S, leaves2, dt)
    full2.backward()
    opt.step()
    ref_p = torch.cat([l.detach().reshape(-1) for l in leaves2]).numpy()
    flat = lambda drv, buf: torch.cat([drv._seg(buf, o, k) for o, k in ((drv.o_means, 3 * n), (drv.o_scales, 3 * n),
                                      (drv.o_opac, n), (drv.o_colors, 3 * S["sh"] * n))]).cpu().numpy()
    step = flat(d, d.p) - flat(d, p0)
    ref_step = ref_p - flat(d, p0).astype(np.float64)
    # The first Adam step is lr*g/(|g|+eps): ~lr*sign(g) wherever |g| >> eps = 1e-8, but hypersensitive to
    # ABSOLUTE gradient error where |g| ~ eps (d step/d g = lr*eps/(|g|+eps)^2 = 2e6 at g = 0).  The gradient
    # contract is relative L2 <= 1e-3 (checked above), so the element-wise bound applies away from g ~ 0.
    full_g = torch.cat([l.grad.reshape(-1) for l in leaves2]).numpy()
    solid = np.abs(full_g) >= 1e-6
    assert solid.mean() > 0.5
    assert np.abs(step - ref_step)[solid].max() <= 2e-3
    assert rel_l2(step, ref_step) <= 2e-2


def test_fit_loop_tracks_oracle_loop():
    S = _setup(4, n=150, V=2, W=40, H=32)
    d = _driver(S)
    dt = torch.float32
    leaves = [torch.nn.Parameter(torch.from_numpy(S[k]).to(dt)) for k in ("means", "scales_raw", "op_raw", "col_raw")]
    opt = torch.optim.Adam(leaves, lr=0.02)
    ours, ref = [], []
    for it in range(25):
        ours.append(float(d.step().item()))
        opt.zero_grad(set_to_none=True)
        data, full = _oracle_loss(S, leaves, dt)
        full.backward()
        opt.step()
        ref.append(float(data))
    assert ref[-1] < ref[0]
    assert abs(ours[0] - ref[0]) <= 1e-5
    assert np.abs(np.array(ours) - np.array(ref)).max() <= 0.02 * ref[0]      # same descent curve
    assert not d.check_overflow()


def test_step_from_host_equals_device_step():
    S = _setup(1)
    d1, d2 = _driver(S), _driver(S)
    host_t = {i: torch.from_numpy(S["tgts"][i]).pin_memory() for i in range(S["V"])}
    host_m = {i: torch.from_numpy(S["masks"][i]).pin_memory() for i in range(S["V"])}
    for _ in range(3):
        l1 = float(d1.step().item())
        l2 = d2.step_from_host(host_t, host_m)
        assert abs(l1 - l2) <= 1e-6
    # same kernels, same inputs: the two runs differ only by the order of the float atomics, which Adam
    # turns into O(lr) differences on the few entries whose gradient is ~eps -- compare in L2
    assert rel_l2(d1.p.cpu().numpy(), d2.p.cpu().numpy()) <= 1e-4


def test_step_from_host_deferred_loss_is_the_same_fit_one_step_late():
    """defer_loss=True returns the PREVIOUS step's loss (None first) and never drains the device; the losses and the
    parameters are those of the blocking loop."""
    S = _setup(1)
    d1, d2 = _driver(S, lanes=2), _driver(S, lanes=2)
    host_t = {i: torch.from_numpy(S["tgts"][i]).pin_memory() for i in range(S["V"])}
    host_m = {i: torch.from_numpy(S["masks"][i]).pin_memory() for i in range(S["V"])}
    blocking = [d1.step_from_host(host_t, host_m) for _ in range(5)]
    deferred = [d2.step_from_host(host_t, host_m, defer_loss=True) for _ in range(5)]
    assert deferred[0] is None
    deferred = deferred[1:] + [d2.flush_loss()]
    assert d2.flush_loss() == deferred[-1]              # nothing pending any more: the last value again
    for a, b in zip(blocking, deferred):
        assert abs(a - b) <= 2e-6
    assert rel_l2(d1.p.cpu().numpy(), d2.p.cpu().numpy()) <= 1e-4
    # a blocking call after deferred ones flushes first and carries on
    l6a, l6b = d1.step_from_host(host_t, host_m), d2.step_from_host(host_t, host_m)
    assert abs(l6a - l6b) <= 2e-6


def test_spatial_reorder_leaves_the_fit_unchanged():
    """FitDriver.reorder_spatial permutes parameters and Adam moments into 3-D Morton order: same losses, and the
    same parameters up to that permutation, as the unpermuted driver."""
    S = _setup(4, V=4)
    d1, d2 = _driver(S), _driver(S)
    for _ in range(2):                      # Adam moments are non-zero when the permutation is applied
        d1.step(); d2.step()
    before_p, before_m = d2.means().clone(), d2.m[d2.o_opac:d2.o_opac + d2.n].clone()
    perm = d2.reorder_spatial()
    assert sorted(perm.cpu().tolist()) == list(range(d2.n))
    assert torch.equal(d2.means(), before_p[perm])                      # parameters ...
    assert torch.equal(d2.m[d2.o_opac:d2.o_opac + d2.n], before_m[perm])   # ... and Adam moments move together
    for _ in range(3):
        l1, l2 = float(d1.step().item()), float(d2.step().item())
        assert abs(l1 - l2) <= 1e-6 * max(1.0, abs(l1))
    for a, b in ((d1.means(), d2.means()), (d1.scales_raw(), d2.scales_raw()), (d1.opacities_raw(), d2.opacities_raw()),
                 (d1.colors_raw(), d2.colors_raw())):
        assert rel_l2(a[perm].cpu().numpy(), b.cpu().numpy()) <= 1e-4
    assert not d2.check_overflow()


def test_checkpoint_resume_continues_the_same_fit(tmp_path):
    """save_checkpoint / load_checkpoint carry raw parameters, Adam moments and the step count: a resumed driver
    takes the same steps as the one that kept running (bit-equal state right after the load)."""
    S = _setup(4, V=3)
    d1 = _driver(S)
    for _ in range(3):
        d1.step()
    ck = tmp_path / "fit.npz"
    d1.save_checkpoint(ck)
    d2 = _driver(S)                       # fresh driver with the initial parameters
    d2.load_checkpoint(ck)
    assert d2.step_no == d1.step_no == 3
    for a, b in ((d1.p, d2.p), (d1.m, d2.m), (d1.v, d2.v)):
        assert torch.equal(a, b)
    for _ in range(2):
        l1, l2 = float(d1.step().item()), float(d2.step().item())
        assert abs(l1 - l2) <= 1e-6 * max(1.0, abs(l1))
    assert rel_l2(d1.p.cpu().numpy(), d2.p.cpu().numpy()) <= 1e-4
    with pytest.raises(ValueError):
        _driver(_setup(1, V=3)).load_checkpoint(ck)


@pytest.mark.parametrize("lanes", [2, 3])
def test_view_lanes_equal_single_stream(lanes):
    """Views spread over concurrent CUDA-stream lanes: same loss (fixed summation order) and the same
    parameters as the one-stream schedule, through both the device-target and the host-target entry."""
    S = _setup(4, V=5)
    d1, d2, d3 = _driver(S), _driver(S, lanes=lanes), _driver(S, lanes=lanes)
    host_t = {i: torch.from_numpy(S["tgts"][i]).pin_memory() for i in range(S["V"])}
    host_m = {i: torch.from_numpy(S["masks"][i]).pin_memory() for i in range(S["V"])}
    for _ in range(3):
        l1 = float(d1.step().item())
        l2 = float(d2.step().item())
        l3 = d3.step_from_host(host_t, host_m)
        assert abs(l1 - l2) <= 1e-6 and abs(l1 - l3) <= 1e-6
    assert rel_l2(d1.p.cpu().numpy(), d2.p.cpu().numpy()) <= 1e-4
    assert rel_l2(d1.p.cpu().numpy(), d3.p.cpu().numpy()) <= 1e-4
    assert not d2.check_overflow()


def test_fused_loss_backward_equals_separate_kernels():
    """b2s_fit_backward_blend (loss evaluated from the accumulators inside the g-buffer kernel) against
    b2s_forward -> b2s_fit_loss -> b2s_backward_blend: same loss, same gradients up to the atomics' order."""
    S = _setup(16, V=4)
    d1, d2 = _driver(S, fused_loss=False), _driver(S, fused_loss=True)
    for _ in range(2):
        l1, l2 = float(d1.step().item()), float(d2.step().item())
        assert abs(l1 - l2) <= 1e-6
        assert rel_l2(d1.g.cpu().numpy(), d2.g.cpu().numpy()) <= 1e-5
    S = _setup(1, V=2)
    for m in (True, False):      # with and without silhouette masks
        d1, d2 = _driver(S, fused_loss=False), _driver(S, fused_loss=True)
        if not m:
            d1.masks, d2.masks = {}, {}
        assert abs(float(d1.step().item()) - float(d2.step().item())) <= 1e-6
        assert rel_l2(d1.g.cpu().numpy(), d2.g.cpu().numpy()) <= 1e-5


@pytest.mark.parametrize("sh", [1, 4, 16])
def test_batched_preprocess_equals_per_view(sh):
    """b2s_preprocess_views + b2s_forward_prepared (parameters read once per iteration) against the per-view
    b2s_forward: identical records, so the same loss and gradients up to the order of the float atomics."""
    S = _setup(sh, V=5)
    d1, d2 = _driver(S, batched_preprocess=False), _driver(S, batched_preprocess=True)
    assert d1.prepared is None and d2.prepared is not None
    for _ in range(2):
        l1, l2 = float(d1.step().item()), float(d2.step().item())
        assert abs(l1 - l2) <= 1e-6        # the loss is summed with float atomics
        assert rel_l2(d1.g.cpu().numpy(), d2.g.cpu().numpy()) <= 1e-5


def test_dropin_training_loop_matches_oracle_loop():
    """The reference script's flow (nn.Parameters -> activations -> per-view render -> loss ->
    backward -> torch Adam) through the drop-in render_gaussians_torch."""
    r = pkg("renderer")
    S = _setup(4, n=120, V=2, W=40, H=32)
    tD = lambda a: torch.from_numpy(a).to(dev())
    P = [torch.nn.Parameter(tD(S[k])) for k in ("means", "scales_raw", "op_raw", "col_raw")]
    opt = torch.optim.Adam(P, lr=0.02)
    cams = [r.Camera(view=tD(v), proj=tD(p)) for v, p in S["cams"]]
    Pc = [torch.nn.Parameter(torch.from_numpy(S[k])) for k in ("means", "scales_raw", "op_raw", "col_raw")]
    optc = torch.optim.Adam(Pc, lr=0.02)
    for it in range(8):
        opt.zero_grad(set_to_none=True)
        sc = torch.nn.functional.softplus(P[1]) + 1e-3
        op = torch.sigmoid(P[2])
        total = torch.tensor(0.0, device=dev())
        for i in range(S["V"]):
            pred, alpha, depth = r.render_gaussians_torch(P[0], sc, P[3], op, cams[i], width=S["W"], height=S["H"],
                                                          background=torch.tensor([0.0, 0.0, 0.0], device=dev()),
                                                          max_gaussians=3000, return_aux=True)
            total = total + r1.fit_loss(pred, alpha, depth, tD(S["tgts"][i]), tD(S["masks"][i]), None)
        loss = total / S["V"] + 1e-3 * op.mean() + 1e-3 * sc.mean()
        loss.backward()
        opt.step()
        optc.zero_grad(set_to_none=True)
        _, full = _oracle_loss(S, Pc, torch.float32)
        full.backward()
        optc.step()
        assert abs(float(loss) - float(full)) <= 2e-4 * max(1.0, abs(float(full)))
    for a, b in zip(P, Pc):
        assert rel_l2(a.detach().cpu().numpy(), b.detach().numpy()) <= 5e-3


def test_launcher_binds_reference_module_names(tmp_path):
    """run_reference_script.py makes `from torch_renderer import ...` / `from device_utils import ...`
    (reference python/fit_multiview_stub.py:12-13) resolve to the CUDA path."""
    script = tmp_path / "mini_fit.py"
    script.write_text(
        "import torch\n"
        "from device_utils import get_default_device\n"
        "from torch_renderer import Camera, look_at, perspective, render_gaussians_torch\n"
        "dev = get_default_device()\n"
        "assert dev.type == 'cuda'\n"
        "proj = perspective(60.0, 1.0, 0.01, 100.0, device=dev)\n"
        "view = look_at(torch.tensor([0.,0.5,2.5], device=dev), torch.zeros(3, device=dev), torch.tensor([0.,1.,0.], device=dev))\n"
        "m = (torch.rand((50,3), device=dev)-0.5).requires_grad_(True)\n"
        "out = render_gaussians_torch(m, torch.full((50,3),0.1,device=dev), torch.rand((50,3),device=dev), torch.full((50,),0.5,device=dev), Camera(view=view, proj=proj), 32, 32)\n"
        "out.mean().backward()\n"
        "assert out.shape == (32,32,3) and m.grad.abs().sum() > 0\n"
        "print('MINI_FIT_OK', render_gaussians_torch.__module__)\n")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "3dgaussian_b200", "run_reference_script.py"), "--seed", "0",
                          str(script)], capture_output=True, text=True, timeout=300)
    assert "MINI_FIT_OK 3dgaussian_b200.renderer" in res.stdout, res.stdout + res.stderr


# ---------------------------------------------------------------- multi GPU -----------------
def _mg_worker(rank, world, port, out, comm="nccl"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    torch.distributed.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    fit = importlib.import_module("3dgaussian_b200.fit")
    S = _setup(4, n=9000, V=5, W=64, H=48)
    cams = [(v.reshape(-1).tolist(), p.reshape(-1).tolist()) for v, p in S["cams"]]
    d = fit.FitDriver(S["n"], S["sh"], S["W"], S["H"], cams, torch.device("cuda", rank), rank=rank, world=world, lanes=2,
                      comm=comm)
    assert len(d._chunks()) == 3          # the pipelined tail: chain rule | coalesced all-reduce | Adam, per chunk
    if comm == "multimem":                # no silent fallback: the fused NVLink-multicast tail is what must run here
        assert d._symm is not None, getattr(d, "comm_fallback", "multimem tail not active")
    t = lambda a: torch.from_numpy(a).to(d.dev)
    d.set_params(t(S["means"]), t(S["scales_raw"]), t(S["op_raw"]), t(S["col_raw"]))
    d.plan()
    d.set_targets({i: t(S["tgts"][i]) for i in d.views}, {i: t(S["masks"][i]) for i in d.views})
    for _ in range(3):
        loss = d.step()
    torch.cuda.synchronize()
    d._sync_moments()                     # multimem: the owners' Adam moments gathered on every rank
    res = {"p": d.p.cpu(), "m": d.m.cpu(), "v": d.v.cpu(), "loss": loss.cpu(), "views": d.views}
    # densify / prune changes the Gaussian count: the flat buffers (symmetric allocations + multicast binding in the
    # multimem mode) are rebuilt on every rank, identically (Philox jitter), and the fit carries on
    k = d.densify_prune(3, max_gaussians=12000, seed=0, reorder=True)
    if comm == "multimem":
        assert d._symm is not None, getattr(d, "comm_fallback", "multimem tail not active after densify")
    for _ in range(2):
        loss2 = d.step()
    torch.cuda.synchronize()
    res.update({"n_after": k, "p_after": d.p.cpu(), "loss_after": loss2.cpu()})
    torch.save(res, out.format(rank))
    torch.distributed.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("comm", ["nccl", "multimem"])
def test_two_gpu_fit_equals_one_gpu(tmp_path, comm):
    """comm="nccl": chunked NCCL all-reduce + replicated Adam; comm="multimem": the fused reduce-scatter + Adam +
    all-gather kernel over NVLink multicast (b2s_adam_step_multimem).  Either way the replicas must be bit-identical
    and equal to the one-GPU fit."""
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "rank{}.pt")
    mp.spawn(_mg_worker, args=(2, port, out, comm), nprocs=2, join=True)
    r0, r1_ = torch.load(out.format(0)), torch.load(out.format(1))
    assert r0["views"] == [0, 2, 4] and r1_["views"] == [1, 3]
    assert torch.equal(r0["p"], r1_["p"])                       # replicas stay bit-identical
    assert torch.equal(r0["m"], r1_["m"]) and torch.equal(r0["v"], r1_["v"])
    assert r0["n_after"] == r1_["n_after"] and r0["n_after"] > 9000
    assert torch.equal(r0["p_after"], r1_["p_after"]) and torch.equal(r0["loss_after"], r1_["loss_after"])
    assert bool(torch.isfinite(r0["p_after"]).all()) and float(r0["loss_after"]) < float(r0["loss"]) * 1.5
    S = _setup(4, n=9000, V=5, W=64, H=48)
    d = _driver(S)
    for _ in range(3):
        loss = d.step()
    assert abs(float(loss.item()) - float(r0["loss"])) <= 1e-6
    assert rel_l2(r0["p"].numpy(), d.p.cpu().numpy()) <= 1e-5
    assert rel_l2(r0["m"].numpy(), d.m.cpu().numpy()) <= 1e-4 and rel_l2(r0["v"].numpy(), d.v.cpu().numpy()) <= 1e-4


_XCHECK = r'''
import importlib, sys, numpy as np, torch
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
import scenes
fit = importlib.import_module("3dgaussian_b200.fit")
dev = torch.device("cuda", 0)
n, V, W, H, sh = 60000, 2, 1920, 1080, 4
means, scales, colors, opac = scenes.make_scene(11, n, sh=sh, s_lo=0.01, s_hi=0.06)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
sr = np.log(np.expm1(np.maximum(scales - 1e-3, 1e-4))).astype(np.float32)
orr = np.log(opac / (1 - opac)).astype(np.float32)
cams = [tuple(m.reshape(-1).tolist() for m in scenes.orbit_camera(i, V, W, H)) for i in range(V)]
d = fit.FitDriver(n, sh, W, H, cams, dev)
d.set_params(t(means), t(sr), t(orr), t(colors)); d.plan()
g = torch.Generator(device="cpu").manual_seed(3)
d.set_targets({{i: torch.rand(H, W, 3, generator=g).to(dev) for i in range(V)}},
              {{i: (torch.rand(H, W, generator=g) > 0.5).float().to(dev) for i in range(V)}})
loss = float(d.step().item())
assert not d.check_overflow()
np.savez({out!r}, loss=np.float64(loss), g=d.g.cpu().numpy())
'''


def test_tcgen05_blend_matches_mma_sync_at_full_hd(tmp_path):
    """The tcgen05 forward / backward blend kernels against the mma.sync ones (B2S_FWD_MMASYNC / B2S_BWD_MMASYNC,
    read once per process: two subprocesses) on a 1080p fit step: 8160 tiles, i.e. work units of up to 4096
    Gaussians, several 128-Gaussian steps per unit, multi-unit tiles, the unit descriptor table."""
    outs = []
    for tag, env in (("umma", {}), ("mma", {"B2S_FWD_MMASYNC": "1", "B2S_BWD_MMASYNC": "1"})):
        out = str(tmp_path / f"{tag}.npz")
        e = dict(os.environ, **env)
        r = subprocess.run([sys.executable, "-c", _XCHECK.format(root=ROOT, out=out)], env=e, capture_output=True, text=True,
                           timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(np.load(out))
    a, b = outs
    assert abs(float(a["loss"]) - float(b["loss"])) <= 5e-6 * max(1.0, abs(float(b["loss"])))
    assert rel_l2(a["g"], b["g"]) <= 1e-3          # the gradient bar of north_star, between two kernel families


# ---------------------------------------------------------------- depth term + overflow guard ----
def _oracle_loss_depth(S, leaves, dt, w_depth=0.05):
    """_oracle_loss with the reference's depth term (fit_multiview_stub.py:298-303) on top."""
    t = lambda a: torch.from_numpy(a).to(dt)
    m, sr, orr, cr = leaves
    sc = torch.nn.functional.softplus(sr) + 1e-3
    op = torch.sigmoid(orr)
    col = torch.sigmoid(cr) if S["sh"] == 1 else cr
    total = torch.zeros((), dtype=dt)
    for i in range(S["V"]):
        rgb, alpha, depth = r1.render_r1(m, sc, col, op, t(S["cams"][i][0]), t(S["cams"][i][1]), S["W"], S["H"])
        total = total + r1.fit_loss(rgb, alpha, depth, t(S["tgts"][i]), t(S["masks"][i]), t(S["depths"][i]),
                                    silhouette_weight=0.2, depth_weight=w_depth)
    return total / S["V"]


@pytest.mark.parametrize("sh,lanes", [(1, 1), (4, 2), (16, 1)])
def test_fit_step_with_depth_term_matches_oracle(sh, lanes):
    """FitDriver(use_depth=True): 5-plane tcgen05 forward (depth plane kept in the accumulators), the per-view depth
    statistics (max, arg-max count, signed sum), the depth term fused into the g-buffer kernel and the depth-gradient
    tcgen05 backward -- loss and all four gradients against autograd of the R1 oracle in float64."""
    capi = pkg("capi")
    S = _setup(sh, n=260, V=3, W=48, H=32, seed=9)
    rng = np.random.RandomState(21)
    S["depths"] = [rng.rand(S["H"], S["W"]).astype(np.float32) for _ in range(S["V"])]
    before = capi.path_counts()
    d = _driver(S, use_depth=True, lanes=lanes)
    t = lambda a: torch.from_numpy(a).to(dev())
    d.set_targets({i: t(S["tgts"][i]) for i in d.views}, {i: t(S["masks"][i]) for i in d.views},
                  {i: t(S["depths"][i]) for i in d.views})
    loss_dev = d.step()
    assert not d.check_overflow()
    after = capi.path_counts()
    assert after["fwd_tcgen05"] - before["fwd_tcgen05"] >= S["V"] and after["bwd_tcgen05"] - before["bwd_tcgen05"] >= S["V"]
    assert after["fwd_other"] == before["fwd_other"] and after["bwd_other"] == before["bwd_other"]
    dt = torch.float64
    leaves = [torch.from_numpy(S[k]).to(dt).requires_grad_(True) for k in ("means", "scales_raw", "op_raw", "col_raw")]
    data = _oracle_loss_depth(S, leaves, dt)
    data.backward()
    ref_g = torch.cat([l.grad.reshape(-1) for l in leaves]).numpy()
    assert abs(float(loss_dev.item()) - float(data)) <= 2e-5
    n = S["n"]
    got = [v.cpu().numpy() for v in d.grad_views()]
    rels = {}
    for name, gv, (a, b) in zip(("means", "scales", "opac", "colors"), got,
                                ((0, 3 * n), (3 * n, 6 * n), (6 * n, 7 * n), (7 * n, None))):
        rels[name] = rel_l2(gv, ref_g[a:b])
    from conftest import report
    report("fit_step_depth_term", sh=sh, lanes=lanes, loss=float(loss_dev.item()), loss_oracle=float(data), **rels)
    for name, v in rels.items():
        assert v <= 1e-3, (name, v)


def test_step_from_host_with_depth_maps_equals_device_step():
    S = _setup(4, n=150, V=3, W=40, H=32)
    rng = np.random.RandomState(2)
    S["depths"] = [rng.rand(S["H"], S["W"]).astype(np.float32) for _ in range(S["V"])]
    t = lambda a: torch.from_numpy(a).to(dev())
    d1, d2 = _driver(S, use_depth=True), _driver(S, use_depth=True, lanes=2)
    d1.set_targets({i: t(S["tgts"][i]) for i in d1.views}, {i: t(S["masks"][i]) for i in d1.views},
                   {i: t(S["depths"][i]) for i in d1.views})
    host = lambda key: {i: torch.from_numpy(S[key][i]).pin_memory() for i in range(S["V"])}
    for _ in range(3):
        l1 = float(d1.step().item())
        l2 = d2.step_from_host(host("tgts"), host("masks"), host("depths"))
        assert abs(l1 - l2) <= 2e-6
    assert rel_l2(d1.p.cpu().numpy(), d2.p.cpu().numpy()) <= 1e-4
    with pytest.raises(ValueError):
        _driver(S).set_targets({}, {}, {0: t(S["depths"][0])})      # depth maps need use_depth=True


def test_overflowing_iteration_never_reaches_the_parameters():
    """Pair buffers sized for small Gaussians, then the scales are blown up: the views overflow, the device guard
    skips Adam (parameters and moments untouched), the driver re-plans and repeats the iteration."""
    S = _setup(1, n=3000, V=2, W=160, H=120, s_lo=0.004, s_hi=0.01)     # ~2 of 80 tiles per Gaussian
    d = _driver(S, overflow_check_every=1000)          # poll only when asked
    d.step()
    assert not d.check_overflow()
    with torch.no_grad():
        d.scales_raw().add_(3.0)                       # sigma x ~20: far more (Gaussian,tile) pairs than planned
    p_before, m_before, step_before = d.p.clone(), d.m.clone(), d.step_no
    d.step()
    torch.cuda.synchronize()
    assert int(d.skipped_dev.item()) == 1
    assert torch.equal(d.p, p_before) and torch.equal(d.m, m_before)      # the guard held
    assert d.check_overflow()                          # re-planned, iteration repeated
    assert d.step_no == step_before + 1 and not torch.equal(d.p, p_before)
    # the repeated step equals a step of a driver planned for the big Gaussians from the start
    S2 = dict(S)
    S2["scales_raw"] = S["scales_raw"].copy()
    d_ref = _driver(S2)
    d_ref.step()
    with torch.no_grad():
        d_ref.scales_raw().add_(3.0)
    d_ref.plan()
    d_ref.step()
    assert not d_ref.check_overflow()
    assert rel_l2(d.p.cpu().numpy(), d_ref.p.cpu().numpy()) <= 1e-4
    # host-fed steps check the guard every step
    with torch.no_grad():
        d.scales_raw().add_(1.5)
    host_t = {i: torch.from_numpy(S["tgts"][i]).pin_memory() for i in range(S["V"])}
    host_m = {i: torch.from_numpy(S["masks"][i]).pin_memory() for i in range(S["V"])}
    n_before = d.step_no
    d.step_from_host(host_t, host_m)
    assert d.step_no == n_before + 1 and int(d.skipped_dev.item()) == 0


@pytest.mark.parametrize("sh,lanes", [(1, 1), (16, 2)])
def test_chunked_tail_equals_single_pass(sh, lanes):
    """The multi-GPU tail (chain rule, all-reduce, Adam pipelined over ranges of Gaussians: b2s_backward_params_range +
    per-slice guarded Adam with the regularisers re-weighted per slice) forced onto one GPU against the single-pass
    tail: same gradients, same parameters after three steps."""
    S = _setup(sh, n=5000, V=3, W=64, H=48)
    d1, d2 = _driver(S, lanes=lanes), _driver(S, lanes=lanes, grad_chunks=3)
    assert len(d2._chunks()) == 2 and len(d1._chunks()) == 1        # 5000 Gaussians: at most ceil(5000/4096) chunks
    S2 = _setup(sh, n=20000, V=2, W=64, H=48)
    d1, d2 = _driver(S2, lanes=lanes), _driver(S2, lanes=lanes, grad_chunks=3)
    assert len(d2._chunks()) == 3
    for _ in range(3):
        l1 = float(d1.step().item())
        l2 = float(d2.step().item())
        assert abs(l1 - l2) <= 1e-6
        for a, b in zip(d1.grad_views(), d2.grad_views()):          # d2's gradient buffer is chunk-major
            assert rel_l2(a.cpu().numpy(), b.cpu().numpy()) <= 1e-5
    assert rel_l2(d1.p.cpu().numpy(), d2.p.cpu().numpy()) <= 1e-5
    assert not d1.check_overflow() and not d2.check_overflow()
