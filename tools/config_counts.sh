#!/bin/bash
# FP64-pipe and total warp instructions of the lean trace kernels of BASELINE configs 2-5 at full size (second call of
# each config: the pure kernels), for bench.py's per-config roofline fractions -> profiles/r02_config_counts.json
ncu --metrics smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none \
    -k regex:trace_lean --csv --log-file gpurun_out/config_counts.csv \
    python examples/run_configs.py --configs 2 3 4 5 --repeat 2 > gpurun_out/config_counts.log 2>&1
python tools/config_counts.py gpurun_out/config_counts.csv > gpurun_out/r02_config_counts.json
cat gpurun_out/r02_config_counts.json
