#!/bin/bash
# A/B the default library against variants built by tools/build_variants.sh:  tools/ab_variants.sh "<profile_trace args>" name1 name2 ...
# (two rounds, so that a drift of the box shows; RTB_LEAN_PURE=2 in the environment runs the pure kernels from launch 1)
args="$1"; shift
for round in 1 2; do
for v in "" "$@"; do
  if [ -n "$v" ]; then export RTB_LIBRARY_PATH=/root/repo/ray_trace_pb_b200/_lib/librtb_$v.so; else unset RTB_LIBRARY_PATH; fi
  echo "== ${v:-base}  $(timeout 120 python tools/profile_trace.py $args 2>&1 | grep best)"
done
done
