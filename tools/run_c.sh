#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest raft+warp"; timeout 900 python -m pytest tests/test_gpu_raft.py tests/test_gpu_warp.py -m gpu -q -s --timeout 600 -p no:cacheprovider > gpurun_out/pytest_c.log 2>&1; echo "pytest rc=$?"; grep -h "EPE\|passed\|failed\|Error" gpurun_out/pytest_c.log | tail -30
echo "=== warp bench"; timeout 300 python tools/warp_bench.py > gpurun_out/warp_bench.log 2>&1; echo "rc=$?"; grep "B=32" gpurun_out/warp_bench.log
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_c.json')); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])"; tail -3 gpurun_out/bench_c.err
