#!/bin/bash
# A/B of the lean kernel's pure instantiations on the relay: RTB_LEAN_PURE=0 (probe-driven kernels only) against the
# default (verdict cache: launch 1 probe-driven, launches 2.. pure).
R=${1:-4e7}
for pure in 0 1; do
  export RTB_LEAN_PURE=$pure
  for m in "--keep last" "--keep none --reduce stats" "--keep last --reduce grid"; do
    echo "== RTB_LEAN_PURE=$pure $m"; timeout 300 python tools/profile_trace.py --rays $R --launches 6 $m 2>&1 | tail -2
  done
done
