#!/bin/bash
# ncu passes (1 GPU): launch list of one eager step; --set full of the hand-written kernels.
mkdir -p gpurun_out
echo "=== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_step.csv python tools/profile_step.py step fp16 > gpurun_out/ncu_step.log 2>&1; echo "rc=$?"
echo "=== full set"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"corr_pyramid_resident|corr_prep16|corr_lookup_kernel|warp_cubic_u8c3|warp_mask_composite" -s 10 -c 5 -o gpurun_out/prof_kernels -f python tools/profile_step.py kernels fp16 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; ls -la gpurun_out/*.ncu-rep
