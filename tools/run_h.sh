#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 1200 python -m pytest tests/test_gpu_blur.py tests/test_gpu_mask.py tests/test_gpu_dropin.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_h.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_h.log
echo "=== warp bench"; timeout 300 python tools/warp_bench.py > gpurun_out/warp_bench.log 2>&1; echo "rc=$?"; grep "composite B=32\|composite B=16" gpurun_out/warp_bench.log
