"""8-GPU all-reduce probe (torchrun): what the 220 MB fp32 gradient exchange of a C4 iteration costs on an otherwise
idle GPU, by NCCL and by torch's symmetric-memory multimem kernels (NVLS: the reduction happens in the NVSwitch).
Prints one line per (method, size) on rank 0: ms, algorithmic GB/s, bus GB/s."""
import os
import sys

import torch
import torch.distributed as dist


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_full = 55_000_064
    sizes = [n_full, n_full // 2, n_full // 4]
    out = []
    buf = torch.ones(n_full, dtype=torch.float32, device="cuda")
    for n in sizes:
        ms = timeit(lambda: dist.all_reduce(buf[:n]))
        out.append(("nccl", n, ms))
    try:
        import torch.distributed._symmetric_memory as symm_mem
        t = symm_mem.empty(n_full, dtype=torch.float32, device=torch.device("cuda", local))
        symm_mem.rendezvous(t, group=dist.group.WORLD)
        gname = dist.group.WORLD.group_name
        t.fill_(1.0)
        for name in ("multimem_all_reduce_", "two_shot_all_reduce_"):
            op = getattr(torch.ops.symm_mem, name)
            for n in sizes:
                try:
                    ms = timeit(lambda: op(t[:n], "sum", gname))
                    out.append((name, n, ms))
                except Exception as e:  # noqa: BLE001
                    out.append((name, n, "error: " + str(e)[:200]))
                    break
        # correctness of the multimem path: ones summed over the ranks
        t.fill_(float(rank + 1))
        torch.ops.symm_mem.multimem_all_reduce_(t, "sum", gname)
        torch.cuda.synchronize()
        expect = world * (world + 1) / 2
        out.append(("multimem_check", n_full, f"min {float(t.min())} max {float(t.max())} expect {expect}"))
    except Exception as e:  # noqa: BLE001
        out.append(("symm_mem", 0, "unavailable: " + str(e)[:300]))
    if rank == 0:
        for name, n, ms in out:
            if isinstance(ms, float):
                gb = n * 4 / 1e9
                print(f"{name:24s} {n * 4 / 1e6:8.1f} MB  {ms:7.3f} ms  alg {gb / ms * 1e3:7.1f} GB/s  bus {gb / ms * 1e3 * 2 * (world - 1) / world:7.1f} GB/s", flush=True)
            else:
                print(f"{name:24s} {ms}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
