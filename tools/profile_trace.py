"""Small driver for ncu: a few launches of the fused trace on the bench workload (no reduction, final slab)."""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402

import bench  # noqa: E402
from ray_trace_pb_b200 import device as dev  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=float, default=2e7)
ap.add_argument("--launches", type=int, default=4)
ap.add_argument("--keep", default="last")
ap.add_argument("--reduce", default="none")
ap.add_argument("--precision", default="f64")
ap.add_argument("--layout", default="rows")
ap.add_argument("--sync", action="store_true", help="synchronise after every launch (the verdict cache then has launch 1's probe counts for launch 2: pure kernels)")
args = ap.parse_args()

system, materials = bench.relay_system()
source, side = bench.beam_source(int(args.rays))
rays = source.generate()
if args.layout == "planes":
    rays = rays.t().contiguous()
reducer = None
if args.reduce != "none":
    reducer = dev.Reducer(12, origin=(8.0, 0, 0), grid_n=2048 if args.reduce == "grid" else 0, half_width=8.0)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.launches + 1)]
ev[0].record()
for i in range(args.launches):
    keep = [int(v) for v in args.keep.split(',')] if args.keep[0].isdigit() or args.keep[0] == '-' else args.keep
    out = dev.trace_tensor(system.surfaces, materials, rays, keep=keep, wavelengths=[bench.WAVELENGTH],
                           reducer=reducer, precision=args.precision, layout=args.layout)
    ev[i + 1].record()
    if args.sync:
        torch.cuda.synchronize()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.launches)]
n = rays.shape[1] if args.layout == "planes" else rays.shape[0]
print(f"rays {n}  ms/launch {['%.3f' % m for m in ms]}  best {n * 10 / min(ms) / 1e-3 / 1e9:.2f} G ray*surf/s")
if out is not None:
    import hashlib
    print("digest", hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()[:16])
