#!/bin/bash
# Build librtb_<name>.so variants for A/B runs on the GPU box (RTB_LIBRARY_PATH selects one).
#   tools/build_variants.sh mb6:-DRTB_TRACE_MIN_BLOCKS=6 u2:"-DRTB_SURFACE_UNROLL2 -DRTB_TRACE_MIN_BLOCKS=7"
# A bare number N is shorthand for mbN:-DRTB_TRACE_MIN_BLOCKS=N.
set -e
cd "$(dirname "$0")/../ray_trace_pb_b200/csrc"
for spec in "$@"; do
  if [[ "$spec" =~ ^[0-9]+$ ]]; then name="mb$spec"; defs="-DRTB_TRACE_MIN_BLOCKS=$spec"; else name="${spec%%:*}"; defs="${spec#*:}"; fi
  mkdir -p ../_lib/var_$name
  for f in rtb_api trace_f64 trace_f64_hinted trace_f64_axial trace_lean trace_fast aux_kernels psf_kernels comm; do
    extra="-fmad=false"; [ $f = trace_fast ] && extra="-fmad=true -prec-div=false -prec-sqrt=false"
    src=$f; [ $f = trace_f64_hinted ] && { src=trace_f64; extra="$extra -DRTB_TU_VARIANT=1"; }
    [ $f = trace_f64_axial ] && { src=trace_f64; extra="$extra -DRTB_TU_VARIANT=2"; }
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo $extra -Xcompiler -fPIC -Xptxas -v \
      $defs -c $src.cu -o ../_lib/var_$name/$f.o 2> ../_lib/var_$name/$f.log &
  done
  wait
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../_lib/librtb_$name.so ../_lib/var_$name/*.o -ldl
  echo "$name: $(grep -A1 'ILb1ELb0ELi0' ../_lib/var_$name/trace_f64.log | grep -E 'spill' | head -1) $(grep -A2 'ILb1ELb0ELi0' ../_lib/var_$name/trace_f64.log | grep Used | head -1 | cut -c1-40)"
done
