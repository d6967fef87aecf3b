#!/bin/bash
# Throughput of the fp64 trace in its output modes on the bench workload (one line each).
R=${1:-4e7}
for m in "--keep last" "--keep 20" "--keep none --reduce stats" "--keep none --reduce grid" "--keep 12 --reduce grid"; do
  echo "== $m"; timeout 300 python tools/profile_trace.py --rays $R --launches 5 $m 2>&1 | tail -2
done
echo "== --keep all (1e7 rays)"; timeout 300 python tools/profile_trace.py --rays 1e7 --launches 5 --keep all 2>&1 | tail -2
