#!/bin/bash
# ncu captures of the pure lean kernels (the verdict cache picks them from the second launch on; profile_trace.py --sync
# waits after every launch, so launch 3's trace kernel -- the sixth lean kernel of the process -- is pure)
set -x
RTB_LEAN_MIN_RAYS=0 ncu --set full --clock-control none --import-source on -k regex:trace_lean -s 5 -c 1 -o gpurun_out/r2_prof_pure_grid -f \
    python tools/profile_trace.py --rays 2e7 --keep last --reduce grid --sync > gpurun_out/r2_ncu_pure_grid.log 2>&1
RTB_LEAN_MIN_RAYS=0 ncu --set full --clock-control none --import-source on -k regex:trace_lean -s 5 -c 1 -o gpurun_out/r2_prof_pure_fast -f \
    python tools/profile_trace.py --rays 2e7 --keep last --sync > gpurun_out/r2_ncu_pure_fast.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trace_lean -s 3 -c 1 -o gpurun_out/r2_prof_pure_opm -f \
    python examples/run_configs.py --configs 4 --repeat 2 > gpurun_out/r2_ncu_pure_opm.log 2>&1
