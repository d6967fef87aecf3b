"""Extended fuzz: random systems beyond the 40 pinned seeds (tests/systems.py: random_system), GPU vs the CPU oracle,
bit for bit, full history.  python tools/fuzz_random_systems.py [first_seed] [n_seeds]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import parity  # noqa: E402
import systems  # noqa: E402
import ray_trace_pb_b200.materials as rtm  # noqa: E402
import ray_trace_pb_b200.raytrace as rt  # noqa: E402
from oracle import oracle  # noqa: E402  (checker only)

first = int(sys.argv[1]) if len(sys.argv) > 1 else 40
count = int(sys.argv[2]) if len(sys.argv) > 2 else 200
bad = 0
hinted = 0
alive = []
for seed in range(first, first + count):
    system, m_in, m_out, rays = systems.random_system(rt, rtm, seed, n_rays=3000)
    want = oracle.ray_trace(system, rays, m_in, m_out, n_threads=8)
    for keep in ("all", "last"):
        got = system.ray_trace(rays, m_in, m_out, keep=keep)
        ref = want if keep == "all" else want[[-1]]
        a, b = parity.canonical(got), parity.canonical(ref)      # all NaNs count as one value, zeros keep their sign
        if not np.array_equal(a, b):
            bad += 1
            print(f"seed {seed} keep={keep}: MISMATCH\n" + parity.mismatch_report(got, ref))
    if isinstance(system.surfaces[0], rt.FlatSurface):
        # the kernels of hinted launches (a hint that does not apply is legal: it only costs speed)
        import torch
        from ray_trace_pb_b200 import device as dev
        mats = [m_in] + system.materials + [m_out]
        got = dev.trace_tensor(system.surfaces, mats, torch.from_numpy(rays).cuda(), keep="all", degenerate_first=True)
        hinted += 1
        if not np.array_equal(parity.canonical(got.cpu().numpy()), parity.canonical(want)):
            bad += 1
            print(f"seed {seed} hinted: MISMATCH\n" + parity.mismatch_report(got.cpu().numpy(), want))
    alive.append(int(np.isfinite(want[-1, :, 0]).sum()))
print(f"seeds {first}..{first + count - 1}: {bad} mismatching traces ({hinted} systems start with a flat and were also "
      f"traced with RTB_HINT_DEGENERATE); rays alive at the end: min {min(alive)}, "
      f"median {int(np.median(alive))}, max {max(alive)} of 3000")
sys.exit(1 if bad else 0)
