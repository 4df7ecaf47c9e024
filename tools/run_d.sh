#!/bin/bash
mkdir -p gpurun_out
echo "=== bench no cudnn benchmark"; timeout 600 python bench.py --steps 20 --warmup 3 --no-cudnn-benchmark --cpu-budget-s 0 > gpurun_out/bench_d_nobench.json 2> gpurun_out/bench_d.err; python -c "
import json; d=json.load(open('gpurun_out/bench_d_nobench.json')); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])"
echo "=== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_step.csv python tools/profile_step.py step fp16 > gpurun_out/ncu_step.log 2>&1; echo "rc=$?"
