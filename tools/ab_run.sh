#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/ab.log
for flags in "" "--no-side-streams" "--cudnn-convf1" "--cudnn-fh2" "--cudnn-convf1 --cudnn-fh2 --no-side-streams" "--no-cudnn-benchmark" "--no-graph"; do
  echo "== $flags" >> gpurun_out/ab.log
  timeout 300 python bench.py --steps 30 --warmup 3 --quick $flags >> gpurun_out/ab.log 2>> gpurun_out/ab.err
done
cat gpurun_out/ab.log | cut -c1-200
