#!/bin/bash
# round-2 ncu captures (run on the GPU box, after the same commands have exited 0 without ncu): launch list of the
# bench command, full captures of the bench-step kernel, the final-slab kernel, the OPM launch and the PSF contraction.
set -x
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/r2_b_plain.json 2> gpurun_out/r2_b_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/r2_ncu_bench.log 2>&1
RTB_LEAN_MIN_RAYS=0 ncu --set full --clock-control none --import-source on -k regex:trace_lean -s 5 -c 1 -o gpurun_out/r2_prof_grid -f \
    python tools/profile_trace.py --rays 2e7 --keep last --reduce grid > gpurun_out/r2_ncu_grid.log 2>&1
RTB_LEAN_MIN_RAYS=0 ncu --set full --clock-control none --import-source on -k regex:trace_lean -s 5 -c 1 -o gpurun_out/r2_prof_fast -f \
    python tools/profile_trace.py --rays 2e7 --keep last > gpurun_out/r2_ncu_fast.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trace_lean -s 1 -c 1 -o gpurun_out/r2_prof_opm -f \
    python examples/run_configs.py --configs 4 > gpurun_out/r2_ncu_opm.log 2>&1
ncu --set full --clock-control none -k regex:zgemm -s 2 -c 2 -o gpurun_out/r2_prof_psf -f python tools/psf_timing.py > gpurun_out/r2_ncu_psf.log 2>&1
