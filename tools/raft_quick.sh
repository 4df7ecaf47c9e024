#!/bin/bash
# one parity test of the fast path + a quick step timing (cheapest check after touching raft_fast.py)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_raft.py -m gpu -q -s --timeout 600 -p no:cacheprovider -k "fast_nhwc or config5 or graph_replay or batched_pairs" > gpurun_out/pytest_raftq.log 2>&1; echo "pytest rc=$?"; grep "fast vs module\|720x1280\|passed\|failed\|Error" gpurun_out/pytest_raftq.log | tail -6
timeout 300 python bench.py --steps 30 --warmup 3 --quick 2> gpurun_out/quick.err | cut -c1-110; tail -2 gpurun_out/quick.err
