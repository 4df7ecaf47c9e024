#!/bin/bash
# RAFT parity tests + one quick step timing
mkdir -p gpurun_out
echo "=== pytest raft"; timeout 1500 python -m pytest tests/test_gpu_raft.py tests/test_gpu_dropin.py -m gpu -q -s --timeout 900 -p no:cacheprovider > gpurun_out/pytest_raft.log 2>&1; echo "pytest rc=$?"; grep "fast vs module\|720x1280\|passed\|failed" gpurun_out/pytest_raft.log | tail -6
timeout 300 python bench.py --steps 30 --warmup 3 --quick 2> gpurun_out/quick.err | cut -c1-110; tail -2 gpurun_out/quick.err
