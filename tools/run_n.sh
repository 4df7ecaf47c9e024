#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest raft"; timeout 1500 python -m pytest tests/test_gpu_raft.py tests/test_gpu_dropin.py -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_n.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_n.log
: > gpurun_out/ab.log
for flags in "" "--steps 40"; do
  echo "== $flags" >> gpurun_out/ab.log
  timeout 300 python bench.py --steps 30 --warmup 3 --quick $flags >> gpurun_out/ab.log 2>> gpurun_out/ab.err
done
cut -c1-120 gpurun_out/ab.log; tail -3 gpurun_out/ab.err
echo "=== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_step.csv python tools/profile_step.py step fp16 > gpurun_out/ncu_step.log 2>&1; echo "rc=$?"
