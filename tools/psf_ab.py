"""A/B of the two PSF contraction kernels (rtb_tune psf_dmma 0 / 1): time per call and agreement of the results."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from ray_trace_pb_b200 import _ffi, device as dev

L = _ffi.lib()
G = 2048
red = dev.Reducer(0, grid_n=G, half_width=3.2)
gen = torch.Generator(device="cuda").manual_seed(1)
red.grid_t.copy_(torch.rand(red.grid_t.shape, generator=gen, device="cuda", dtype=torch.float64) * 2 - 1)
red.grid_t[2].fill_(1.0)
for M in (257, 513, 1025):
    out = {}
    for mode in (0, 1):
        L.rtb_tune(b"psf_dmma", mode)
        best = 1e9
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); psf, field = red.psf(M, 0.01, field=True); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        out[mode] = (best, field)
        flop = 8.0 * M * G * G + 8.0 * M * M * G
        print(f"G={G} M={M} {'DMMA' if mode else 'SIMT'}: {best:7.3f} ms  {flop / best / 1e9:6.2f} TFLOP/s fp64")
    d = (out[0][1] - out[1][1]).abs().max().item() / out[0][1].abs().max().item()
    print(f"   max |E_simt - E_dmma| / max |E| = {d:.2e}")
L.rtb_tune(b"psf_dmma", 1)
