#!/bin/bash
# usage: prof_cfg.sh <config>   : full ncu capture of the longest trace_f64 launch of examples/run_configs.py --configs N
cfg=$1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:trace_f64 --csv --log-file gpurun_out/cfg${cfg}_launches.csv python examples/run_configs.py --configs $cfg > /dev/null 2>&1
idx=$(python - <<PY
import csv
rows=list(csv.reader(l for l in open('gpurun_out/cfg${cfg}_launches.csv') if not l.startswith('==')))
h=rows[0]; iv=h.index('Metric Value'); iu=h.index('Metric Unit')
def ms(r):
    v=float(r[iv].replace(',','')); u=r[iu]
    return v/1e6 if u=='ns' else v/1e3 if u=='us' else v*1e3 if u=='s' else v
t=[ms(r) for r in rows[1:] if len(r)>iv]
print(max(range(len(t)), key=lambda i:t[i]))
PY
)
echo "longest trace launch index: $idx"
ncu --set full --clock-control none --import-source on -k regex:trace_f64 -s $idx -c 1 -o gpurun_out/prof_cfg${cfg} -f python examples/run_configs.py --configs $cfg > gpurun_out/ncu_cfg${cfg}.log 2>&1
