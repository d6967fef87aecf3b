"""What a chunked host->device->host pipeline can move on this box, without any kernel: chunk c goes H2D then D2H on
stream c % k (pinned buffers, 1 GiB each way, like the e2e leg of bench.py).  The ceiling for rtb_trace_host."""
import sys
import time

import torch

n_bytes = 1 << 30
h_in = torch.empty(n_bytes, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n_bytes, dtype=torch.uint8).pin_memory()
h_in.fill_(1)


def run(chunk_mib, k, reps=5):
    chunk = chunk_mib << 20
    dev = [torch.empty(chunk, dtype=torch.uint8, device="cuda") for _ in range(k)]
    streams = [torch.cuda.Stream() for _ in range(k)]
    best = None
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for c in range(n_bytes // chunk):
            s = streams[c % k]
            with torch.cuda.stream(s):
                dev[c % k].copy_(h_in[c * chunk:(c + 1) * chunk], non_blocking=True)
                h_out[c * chunk:(c + 1) * chunk].copy_(dev[c % k], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best


for chunk_mib in (8, 16, 32, 64, 128):
    for k in (2, 3, 4, 6):
        dt = run(chunk_mib, k)
        print(f"chunk {chunk_mib:4d} MiB  {k} streams  {dt * 1e3:7.2f} ms  {2 * n_bytes / dt / 1e9:6.1f} GB/s", flush=True)
