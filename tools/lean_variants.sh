#!/bin/bash
# lean kernel under library variants (tools/build_variants.sh): tools/lean_variants.sh "<rays>" name1 name2 ...
R=$1; shift
export RTB_LEAN_MIN_RAYS=0
for v in "" "$@"; do
  if [ -n "$v" ]; then export RTB_LIBRARY_PATH=/root/repo/ray_trace_pb_b200/_lib/librtb_$v.so; fi
  for m in "--keep last" "--keep last --reduce grid"; do
    echo "== ${v:-base} $m"; timeout 120 python tools/profile_trace.py --rays $R --launches 5 $m 2>&1 | tail -2 | head -1
  done
done
