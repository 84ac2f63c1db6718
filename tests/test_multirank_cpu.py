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
