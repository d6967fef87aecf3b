#!/bin/bash
# round-2 final ncu captures (run on the GPU box, each after the same command has exited 0 without ncu):
#   launch list of the bench command; full captures of the PURE lean kernels (bench step, final slab, the OPM launch).
# profile_trace.py --sync waits after every launch, so the verdict cache has launch 1's probe counts for launch 2 and
# the sixth lean kernel of the process (launch 3's trace) is a pure instantiation.
set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2f_b_plain.json 2> gpurun_out/r2f_b_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2f_ncu_bench.log 2>&1
python tools/profile_trace.py --rays 2e7 --keep last --reduce grid --sync > gpurun_out/r2f_plain_grid.log 2>&1 || exit 1
RTB_LEAN_MIN_RAYS=0 ncu --set full --clock-control none --import-source on -k regex:trace_lean -s 5 -c 1 -o gpurun_out/r2f_prof_grid -f \
    python tools/profile_trace.py --rays 2e7 --keep last --reduce grid --sync > gpurun_out/r2f_ncu_grid.log 2>&1
RTB_LEAN_MIN_RAYS=0 ncu --set full --clock-control none --import-source on -k regex:trace_lean -s 5 -c 1 -o gpurun_out/r2f_prof_fast -f \
    python tools/profile_trace.py --rays 2e7 --keep last --sync > gpurun_out/r2f_ncu_fast.log 2>&1
python examples/run_configs.py --configs 4 > gpurun_out/r2f_plain_opm.log 2>&1 || exit 1
# config 4 = two launches (2^27 + the rest) per call, each a probe + a trace; call 1 is probe-driven, call 2 pure
ncu --set full --clock-control none --import-source on -k regex:trace_lean -s 5 -c 1 -o /tmp/r2f_prof_opm -f \
    python examples/run_configs.py --configs 4 --repeat 2 > gpurun_out/r2f_ncu_opm.log 2>&1
# (gpurun brings back at most 64 MiB: the OPM capture stays on the box, its summary and opcode mix travel)
python tools/ncu_summary.py /tmp/r2f_prof_opm.ncu-rep > gpurun_out/r2f_opm_summary.txt 2>&1
python tools/ncu_opmix.py /tmp/r2f_prof_opm.ncu-rep 4194304 >> gpurun_out/r2f_opm_summary.txt 2>&1
