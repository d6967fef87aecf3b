#!/bin/bash
# Install the UNMODIFIED reference (QI2lab/ray_trace_pb, pure Python) into baseline/_ref/ so that bench.py can time its
# NumPy path -- System.ray_trace, raytrace.py:641-661 -- on the GPU box's host cores in the same run as the GPU numbers.
# baseline/_ref/ is git-ignored (reference sources never enter this repo's history) but NOT gpurun-ignored, so it
# travels to the GPU box with the snapshot.  Run in the build container (needs /root/reference); __graft_entry__.build()
# calls it when /root/reference exists.  The source tree is read-only, so the wheel is built from a copy under /tmp;
# --no-deps because matplotlib (a declared dependency the hot path never touches) is not in the wheelhouse.
set -e
REF=${1:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
[ -d "$REF/src/raytrace" ] || { echo "no reference at $REF"; exit 1; }
TMP=$(mktemp -d)
cp -r "$REF" "$TMP/ref"
rm -rf "$HERE/_ref"
python -m pip install --quiet --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$HERE/_ref" "$TMP/ref"
rm -rf "$TMP"
cmp "$REF/src/raytrace/raytrace.py" "$HERE/_ref/raytrace/raytrace.py"
cmp "$REF/src/raytrace/materials.py" "$HERE/_ref/raytrace/materials.py"
echo "reference installed unmodified in $HERE/_ref"
