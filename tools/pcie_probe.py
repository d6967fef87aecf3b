"""Host<->device copy bandwidth with pinned memory (the ceiling of the host-buffer path)."""
import time
import torch

n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


t = timed(lambda: d_a.copy_(h_in, non_blocking=True))
print(f"H2D 1 GiB: {n / t / 1e9:.1f} GB/s")
t = timed(lambda: h_out.copy_(d_b, non_blocking=True))
print(f"D2H 1 GiB: {n / t / 1e9:.1f} GB/s")


def both():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


t = timed(both)
print(f"H2D + D2H concurrently, 1 GiB each: {2 * n / t / 1e9:.1f} GB/s aggregate")
t0 = time.perf_counter()
x = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
print(f"pin 1 GiB: {time.perf_counter() - t0:.3f} s")
