#!/bin/bash
# full GPU check: smoke, all GPU tests, diag, bench (graph), reference arm
mkdir -p gpurun_out
echo "=== smoke"; timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log; grep -h "EPE" gpurun_out/pytest.log | head -20
echo "=== diag"; timeout 900 python tools/gpu_diag.py > gpurun_out/diag_stdout.log 2>&1; grep -v "^(" gpurun_out/diag.txt | tail -32
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; echo "rc=$?"; cat gpurun_out/bench_graph.json; tail -3 gpurun_out/bench_graph.err
