#!/bin/bash
# examples/run_configs.py under the default library and under variants:  tools/ab_configs.sh "4 5" name1 name2
cfgs="$1"; shift
for v in "" "$@"; do
  if [ -n "$v" ]; then export RTB_LIBRARY_PATH=/root/repo/ray_trace_pb_b200/_lib/librtb_$v.so; fi
  echo "== ${v:-base}"; timeout 300 python examples/run_configs.py --configs $cfgs 2>&1 | grep -E "ms \(|ms;" | cut -c1-120
done
