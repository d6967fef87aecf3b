#!/bin/bash
# A/B the default library against variants built by tools/build_variants.sh:  tools/ab_variants.sh "<profile_trace args>" name1 name2 ...
args="$1"; shift
for v in "" "$@"; do
  if [ -n "$v" ]; then export RTB_LIBRARY_PATH=/root/repo/ray_trace_pb_b200/_lib/librtb_$v.so; fi
  echo "== ${v:-base}"; timeout 120 python tools/profile_trace.py $args 2>&1 | tail -2
done
