#!/bin/bash
# round-2 evidence on one B200: smoke, all GPU tests, sanitizers, bench (both arms), launch list of the step, full-set ncu of
# the hand-written kernels.  Everything lands in gpurun_out/r2f_*; the .ncu-rep stays on the box (64 MiB merge limit), only
# its text summary comes back.
mkdir -p gpurun_out
echo "=== smoke"; timeout 600 python __graft_entry__.py smoke > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2f_smoke.log
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2f_pytest.log
echo "=== bench reference"; timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err; echo "rc=$?"; cut -c1-200 gpurun_out/r2f_bench_reference.json
echo "=== bench ours"; timeout 900 python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2f_bench_n1.json')); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'],'clocks',d['clocks']); print('roofline',d['roofline']['frac'],d['roofline'].get('us_per_launch')); [print(' ',x['kernel'][:60],round(x['frac'],3),x.get('us_per_launch')) for x in d['roofline_extra']]; print('batched',d['batched'].get('value'),'clip',d['clip']['value'],'config4',d['config4']['value'],'config5',d['config5']['value']); print('cpu',d['cpu_baseline']['value'],d['cpu_baseline']['cores']); print('reference_gpu',json.dumps(d.get('reference_gpu'))[:600])"; tail -3 gpurun_out/r2f_bench_n1.err
echo "=== launch list (warm caches)"
timeout 900 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2f_launches_step.csv python tools/profile_step.py step fp16 > gpurun_out/r2f_ncu_step.log 2>&1; echo "rc=$?"; wc -l gpurun_out/r2f_launches_step.csv
echo "=== full set"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"corr_pyramid_resident|corr_prep16|corr_absmax|corr_lookup_kernel|warp_cubic_u8c3|warp_mask_composite|instnorm_stats|instnorm_apply|conv7x7|flowhead2_taps|flowhead2_gather|blur_composite|motion_tail16_h|gru_rh_h|gru_update_h|flow_im2col7" -s 46 -c 23 -o /tmp/r2f_kernels -f python tools/profile_step.py kernels fp16 > gpurun_out/r2f_ncu_full.log 2>&1; echo "rc=$?"
python tools/summarize_ncu.py /tmp/r2f_kernels.ncu-rep "ncu --set full, hand-written kernels at the batch-1 / 32-frame sizes (round 2, final)" > gpurun_out/r2f_kernels_ncu_summary.txt 2>&1; wc -l gpurun_out/r2f_kernels_ncu_summary.txt
echo "=== sanitizers"; tools/run_sanitizer.sh gpurun_out > gpurun_out/r2f_san.log 2>&1; tail -12 gpurun_out/r2f_san.log; rm -f gpurun_out/sanitize_driver
