#!/bin/bash
mkdir -p gpurun_out
echo "=== smoke"; timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "=== pytest"; timeout 1200 python -m pytest tests/test_gpu_mask.py tests/test_gpu_dropin.py tests/test_gpu_warp.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_g.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_g.log
echo "=== warp bench"; timeout 300 python tools/warp_bench.py > gpurun_out/warp_bench.log 2>&1; echo "rc=$?"; grep "B=32\|B= 1" gpurun_out/warp_bench.log
