#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_i.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_i.log
echo "=== warp bench"; timeout 300 python tools/warp_bench.py > gpurun_out/warp_bench.log 2>&1; echo "rc=$?"; grep "B=32" gpurun_out/warp_bench.log
