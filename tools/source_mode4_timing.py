"""Final slab / nothing + one reduction with rays generated in the kernel (relay, 4e7 rays): the MODE 4 kernels."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import torch, bench
from ray_trace_pb_b200 import device as dev
system, materials = bench.relay_system()
source, side = bench.beam_source(int(4e7))
for keep in ("last","none"):
    red = dev.Reducer(12, origin=(8.0,0,0), grid_n=2048, half_width=8.0)
    packed = dev.prepare(system.surfaces, materials, [bench.WAVELENGTH])
    ms=[]
    for i in range(5):
        red.reset()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); dev.trace_source(system.surfaces, materials, source, keep=keep, reducer=red, packed=packed); e1.record()
        torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    print(f"source-fused keep={keep}: {source.n_rays*10/min(ms)/1e6:.2f} G ray*surf/s  {['%.3f'%m for m in ms]}")
