#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 1200 python -m pytest tests/test_gpu_raft.py tests/test_gpu_dropin.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_k.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_k.log
: > gpurun_out/ab.log
for flags in "" ; do
  echo "== $flags" >> gpurun_out/ab.log
  timeout 300 python bench.py --steps 30 --warmup 3 --quick $flags >> gpurun_out/ab.log 2>> gpurun_out/ab.err
done
cut -c1-120 gpurun_out/ab.log; tail -3 gpurun_out/ab.err
