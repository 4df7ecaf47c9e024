#!/bin/bash
# warp-family check: parity tests, CUDA-event timings over flow types (optionally one ncu capture: NCU=1)
mkdir -p gpurun_out
echo "=== pytest"; timeout 600 python -m pytest tests/test_gpu_warp.py tests/test_gpu_keyframe.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_warp.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_warp.log
echo "=== warp bench"; timeout 300 python tools/warp_bench.py > gpurun_out/warp_bench.log 2>&1; echo "rc=$?"; grep "B=32\|detect" gpurun_out/warp_bench.log
if [ -n "$NCU" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:warp_cubic_u8c3_tiled -s 50 -c 1 -o gpurun_out/prof_warp_zero -f python tools/warp_bench.py > gpurun_out/ncu_warp.log 2>&1; echo "ncu rc=$?"
fi
