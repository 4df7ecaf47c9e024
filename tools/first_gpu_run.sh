#!/bin/bash
# One gpurun call: smoke, diagnostics, GPU tests, bench (eager + graph), ncu launch list.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
echo "=== smoke"; timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "=== diag"; timeout 900 python tools/gpu_diag.py > gpurun_out/diag_stdout.log 2>&1; echo "diag rc=$?"; tail -60 gpurun_out/diag.txt
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest.log
echo "=== bench eager"; timeout 600 python bench.py --no-graph --steps 5 --warmup 3 --cpu-budget-s 0 > gpurun_out/bench_eager.json 2> gpurun_out/bench_eager.err; echo "rc=$?"; cat gpurun_out/bench_eager.json; tail -5 gpurun_out/bench_eager.err
echo "=== bench graph"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; echo "rc=$?"; cat gpurun_out/bench_graph.json; tail -5 gpurun_out/bench_graph.err
echo "=== bench reference"; timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
