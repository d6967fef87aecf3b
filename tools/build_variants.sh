#!/bin/bash
# Build librtb_mb<N>.so variants (RTB_TRACE_MIN_BLOCKS = N) for A/B runs on the GPU box (RTB_LIBRARY_PATH selects one).
set -e
cd "$(dirname "$0")/../ray_trace_pb_b200/csrc"
for mb in "$@"; do
  mkdir -p ../_lib/var$mb
  for f in rtb_api trace_f64 trace_f32 aux_kernels psf_kernels; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -Xptxas -v \
      -DRTB_TRACE_MIN_BLOCKS=$mb -c $f.cu -o ../_lib/var$mb/$f.o 2> ../_lib/var$mb/$f.log &
  done
  wait
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../_lib/librtb_mb$mb.so ../_lib/var$mb/*.o
  echo "mb=$mb: $(grep -A1 'ILb1ELb0ELi0' ../_lib/var$mb/trace_f64.log | grep -E 'spill' | head -1) $(grep 'Used' ../_lib/var$mb/trace_f64.log | sed -n 7p | cut -c1-40)"
done
