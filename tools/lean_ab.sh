#!/bin/bash
# A/B of the lean kernel against the round-1 kernels on the bench workload: final slab, statistics, grid.
R=${1:-4e7}
for lean in -1 0; do
  export RTB_LEAN_MIN_RAYS=$lean
  for m in "--keep last" "--keep none --reduce stats" "--keep last --reduce grid"; do
    echo "== RTB_LEAN_MIN_RAYS=$lean $m"; timeout 300 python tools/profile_trace.py --rays $R --launches 5 $m 2>&1 | tail -2
  done
done
