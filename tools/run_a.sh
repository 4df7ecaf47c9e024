#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest warp+ref"; timeout 600 python -m pytest tests/test_gpu_warp.py tests/test_gpu_ref_kernel.py tests/test_gpu_mask.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_a.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_a.log
echo "=== warp bench"; timeout 300 python tools/warp_bench.py > gpurun_out/warp_bench.log 2>&1; echo "rc=$?"; cat gpurun_out/warp_bench.log
