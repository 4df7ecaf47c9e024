#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 1200 python -m pytest tests/test_gpu_blur.py tests/test_gpu_keyframe.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_l.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_l.log
echo "=== warp bench"; timeout 300 python tools/warp_bench.py > gpurun_out/warp_bench.log 2>&1; echo "rc=$?"; grep "blur\|resize\|detect" gpurun_out/warp_bench.log; tail -3 gpurun_out/warp_bench.log
