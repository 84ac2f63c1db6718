"""World-size-2 gloo test (CPU) of the host logic of the multi-GPU fit: round-robin view
sharding (fit.local_views) + ONE sum all-reduce of the flat gradient buffer reproduces the
single-process gradient of  sum_i loss_i / V  (reference python/fit_multiview_stub.py:277-308).
Per-view gradients come from the CPU oracle here; on the GPU box the same comparison runs
with the CUDA path (tests/test_gpu_fit.py, bench.py --gpus N)."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import scenes
from oracle import r1_oracle as r1

V, N, W, H = 5, 40, 24, 16


def _flat_grad(view_ids):
    means, scales, colors, opac = scenes.make_scene(31, N, sh=4, s_lo=0.05, s_hi=0.2)
    t = lambda a: torch.from_numpy(a).double()
    leaves = [t(a).requires_grad_(True) for a in (means, scales, opac, colors)]
    total = torch.zeros((), dtype=torch.float64)
    for i in view_ids:
        view, proj = scenes.orbit_camera(i, V, W, H)
        rgb, alpha, depth = r1.render_r1(leaves[0], leaves[1], leaves[3], leaves[2], t(view), t(proj), W, H)
        tgt = torch.rand((H, W, 3), generator=torch.Generator().manual_seed(100 + i)).double()
        total = total + r1.fit_loss(rgb, alpha, depth, tgt, None, None) / V
    total.backward()
    return torch.cat([l.grad.reshape(-1) for l in leaves]), float(total)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fit = importlib.import_module("3dgaussian_b200.fit")
    mine = fit.local_views(V, rank, world)
    g, loss = _flat_grad(mine)
    lt = torch.tensor([loss], dtype=torch.float64)
    dist.all_reduce(g)
    dist.all_reduce(lt)
    if rank == 0:
        torch.save({"g": g, "loss": lt, "mine": mine}, out)
    dist.destroy_process_group()


def test_local_views_partition():
    fit = importlib.import_module("3dgaussian_b200.fit")
    for world in (1, 2, 3, 8):
        parts = [fit.local_views(64, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(64))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
    assert fit.local_views(3, 5, 8) == []          # more ranks than views: idle rank


def test_chunks_and_chunk_major_layout():
    """Host arithmetic of the pipelined tail: the Gaussian ranges tile [0, n) with 64-aligned starts, and the chunk-major
    gradient buffer places every (chunk, segment) slice 64-float aligned, in order, without overlap."""
    fit = importlib.import_module("3dgaussian_b200.fit")
    for n in (1, 100, 4096, 9000, 250_001, 1_000_000):
        for c in (1, 2, 3, 4, 7):
            ch = fit.gaussian_chunks(n, c)
            assert ch[0][0] == 0 and sum(k for _, k in ch) == n and len(ch) <= c
            assert all(f % 64 == 0 for f, _ in ch) and all(a + k == b for (a, k), (b, _) in zip(ch, ch[1:]))
            for sh in (1, 4, 16):
                offs, total = fit.chunk_major_offsets(ch, sh)
                pos = 64
                for (first, cnt), segs in zip(ch, offs):
                    for q, k in enumerate((3, 3, 1, 3 * sh)):
                        assert segs[q] == pos and segs[q] % 64 == 0
                        pos = segs[q] + (k * cnt + 63) // 64 * 64
                    assert segs[4] == pos
                assert total == pos


def _owner_worker(rank, world, port, out):
    """CPU emulation of the fused multi-GPU tail (b2s_adam_step_multimem): every rank holds its own gradient; the owner
    of a share takes the SUM of that share (on the GPU: multimem.ld_reduce through the NVSwitch; here an all-reduce read on
    the share only), applies Adam with ITS moments, and the new parameters of all shares are gathered on every rank
    (multimem.st; here a sum of share-masked buffers)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    capi = importlib.import_module("3dgaussian_b200.capi")
    count = 1003
    gen = torch.Generator().manual_seed(7)
    p0 = torch.randn(count, generator=gen, dtype=torch.float64)
    grads = [torch.randn(count, generator=gen, dtype=torch.float64) for _ in range(world)]
    p, m, v = p0.clone(), torch.zeros(count, dtype=torch.float64), torch.zeros(count, dtype=torch.float64)
    for step in (1, 2, 3):
        g = grads[rank] * step
        red = g.clone()
        dist.all_reduce(red)                                   # what ld_reduce returns for any address
        lo, hi = capi.multimem_share(count, rank, world)
        newp = torch.zeros(count, dtype=torch.float64)
        pn, mn, vn = r1.adam_step(p[lo:hi], red[lo:hi], m[lo:hi], v[lo:hi], step, 0.02)
        m[lo:hi], v[lo:hi] = mn, vn                            # moments live on the owner only
        newp[lo:hi] = pn
        dist.all_reduce(newp)                                  # every share lands on every rank
        p = newp
    torch.save({"p": p, "m": m, "v": v, "share": capi.multimem_share(count, rank, world)}, out.format(rank))
    dist.destroy_process_group()


def test_owner_update_scheme_equals_replicated_adam(tmp_path):
    world, count = 2, 1003
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "own{}.pt")
    mp.spawn(_owner_worker, args=(world, port, out), nprocs=world, join=True)
    res = [torch.load(out.format(r)) for r in range(world)]
    assert torch.equal(res[0]["p"], res[1]["p"])                # replicas identical by construction
    gen = torch.Generator().manual_seed(7)
    p = torch.randn(count, generator=gen, dtype=torch.float64)
    grads = [torch.randn(count, generator=gen, dtype=torch.float64) for _ in range(world)]
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in (1, 2, 3):
        p, m, v = r1.adam_step(p, sum(grads) * step, m, v, step, 0.02)
    assert torch.allclose(res[0]["p"], p, rtol=1e-12, atol=1e-14)
    for r in range(world):                                      # each owner's moments are the replicated ones on its share
        lo, hi = res[r]["share"]
        assert torch.allclose(res[r]["m"][lo:hi], m[lo:hi], rtol=1e-12, atol=1e-15)
        assert torch.allclose(res[r]["v"][lo:hi], v[lo:hi], rtol=1e-12, atol=1e-15)


def test_two_rank_allreduce_equals_single_process(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    assert got["mine"] == [0, 2, 4]
    ref_g, ref_loss = _flat_grad(range(V))
    assert abs(float(got["loss"]) - ref_loss) <= 1e-12
    assert torch.allclose(got["g"], ref_g, rtol=1e-10, atol=1e-14)
