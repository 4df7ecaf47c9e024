#!/bin/bash
# all GPU tests + one bench line (no ncu)
mkdir -p gpurun_out
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest.log
echo "=== bench"; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'],'clocks',d['clocks']); print('roofline',d['roofline']['frac'],d['roofline']['us_per_launch']); [print(' ',x['kernel'][:60],round(x['frac'],3),x.get('us_per_launch')) for x in d['roofline_extra']]; print('batched',d['batched']['value'],'clip',d['clip']['value']); print('cpu',d['cpu_baseline']['value'],d['cpu_baseline']['cores'])"; tail -3 gpurun_out/bench_n1.err
