"""PCIe rates of the box with pinned host memory: H2D alone, D2H alone, both directions at once (what bounds the
host-to-host round trip: 604 MB in, 654 MB out per 64 images)."""
import time
import torch

n = 604 * 1024 * 1024
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def best(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t = 1e9
    for _ in range(reps):
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


t = best(h2d)
print(f"H2D alone  {n / t / 1e9:6.1f} GB/s")
t = best(d2h)
print(f"D2H alone  {n / t / 1e9:6.1f} GB/s")
t = best(both)
print(f"both       {n / t / 1e9:6.1f} GB/s per direction ({2 * n / t / 1e9:.1f} GB/s total)")
