#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_f.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_f.log
: > gpurun_out/ab.log
for flags in "" "--cudnn-fh2"; do
  echo "== $flags" >> gpurun_out/ab.log
  timeout 300 python bench.py --steps 30 --warmup 3 --quick $flags >> gpurun_out/ab.log 2>> gpurun_out/ab.err
done
cut -c1-120 gpurun_out/ab.log
