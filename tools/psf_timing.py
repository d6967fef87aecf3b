"""PSF contraction timing: 2048^2 pupil grid -> M x M samples (CUDA events), with the FP64 rate it corresponds to."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from ray_trace_pb_b200 import device as dev

G = 2048
for M in (257, 513, 1025):
    red = dev.Reducer(0, grid_n=G, half_width=3.2)
    red.grid_t[0].uniform_(-1, 1); red.grid_t[1].uniform_(-1, 1); red.grid_t[2].fill_(1.0)
    for norm in (False, True):
        best = 1e9
        for _ in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); red.psf(M, 0.01, normalize_by_count=norm); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        flop = 8.0 * M * G * G + 8.0 * M * M * G
        print(f"G={G} M={M} normalize={norm}: {best:7.3f} ms  {flop / best / 1e9:6.2f} TFLOP/s fp64 (complex MACs as 8 flops)")
