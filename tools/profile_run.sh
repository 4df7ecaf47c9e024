#!/bin/bash
# ncu passes (1 GPU): launch list of one eager step; --set full of the hand-written kernels.
mkdir -p gpurun_out
echo "=== pytest"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest.log
echo "=== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_step.csv python tools/profile_step.py step > gpurun_out/ncu_step.log 2>&1; echo "rc=$?"
echo "=== full set"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"corr_volume_tc|corr_lookup_kernel|warp_cubic_u8c3|warp_mask_composite" -s 8 -c 4 -o gpurun_out/prof_kernels -f python tools/profile_step.py kernels > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; ls -la gpurun_out/*.ncu-rep
