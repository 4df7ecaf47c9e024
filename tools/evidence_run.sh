#!/bin/bash
# round-1 evidence: bench (both arms), launch list, full-set ncu of the hand-written kernels
mkdir -p gpurun_out
echo "=== bench reference"; timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_ref.json
echo "=== bench ours"; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'],'clocks',d['clocks']); print('roofline',d['roofline']['frac'],d['roofline']['us_per_launch']); [print(' ',x['kernel'][:60],round(x['frac'],3),x.get('us_per_launch')) for x in d['roofline_extra']]; print('batched',d['batched']); print('cpu',d['cpu_baseline']['value'],d['cpu_baseline']['cores'])"; tail -3 gpurun_out/bench_n1.err
echo "=== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_step.csv python tools/profile_step.py step fp16 > gpurun_out/ncu_step.log 2>&1; echo "rc=$?"
echo "=== full set"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"corr_pyramid_resident|corr_prep16|corr_lookup_kernel|warp_cubic_u8c3|warp_mask_composite|instnorm_stats|instnorm_apply|conv7x7|flowhead2_taps|blur_composite" -s 11 -c 11 -o gpurun_out/prof_kernels -f python tools/profile_step.py kernels fp16 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; ls -la gpurun_out/*.ncu-rep
