"""PCIe rates of the box with pinned host memory: H2D alone, D2H alone, both directions at once (what bounds the
host-to-host round trip: 604 MB in, 654 MB out per 64 images).  Under torchrun every rank drives its own GPU AT THE
SAME TIME (barrier before every timed repetition) and rank 0 prints the per-rank mean and the aggregate over the box:
the bound of the N-GPU end-to-end numbers, whose ranks share the host's memory and PCIe fabric."""
import os
import time

import torch

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))  # like bench.py
    except Exception:
        pass
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

n = 604 * 1024 * 1024
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def best(fn, reps=5):
    fn()
    barrier()
    t = 1e9
    for _ in range(reps):
        barrier()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        t = min(t, time.perf_counter() - t0)
    return t


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


def report(name, t, factor=1):
    rate = torch.tensor([factor * n / t / 1e9], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(rate)
    if rank == 0:
        print(f"{name:10s} {rate.item() / world:6.1f} GB/s per GPU, {rate.item():7.1f} GB/s over {world} GPU(s) at once")


report("H2D alone", best(h2d))
report("D2H alone", best(d2h))
report("both", best(both), 2)
if world > 1:
    dist.destroy_process_group()
