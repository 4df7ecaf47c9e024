#!/bin/bash
# compute-sanitizer over the hand-written kernels, through the C ABI, with a standalone driver (no Python in the process).
#   tools/run_sanitizer.sh [outdir]      (GPU box; ~2-4 minutes)
# Writes <outdir>/sanitizer_{memcheck,racecheck,synccheck}.log; each log ends with the tool's summary line and "rc=<exit code>".
set -u
cd "$(dirname "$0")/.."
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
LIB=sd_animation_optical_flow_b200/lib
[ -f $LIB/libsdof_b200.so ] || python -m sd_animation_optical_flow_b200.build
nvcc -O1 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Iinclude tools/sanitize/driver.cu \
     -L$LIB -lsdof_b200 -Xlinker -rpath -Xlinker "$PWD/$LIB" -o "$OUT/sanitize_driver" || exit 1
"$OUT/sanitize_driver" all > "$OUT/sanitizer_plain.log" 2>&1; echo "rc=$?" >> "$OUT/sanitizer_plain.log"
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 "$OUT/sanitize_driver" all > "$OUT/sanitizer_$tool.log" 2>&1
  echo "rc=$?" >> "$OUT/sanitizer_$tool.log"
  tail -4 "$OUT/sanitizer_$tool.log"
done
