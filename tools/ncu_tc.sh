#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"corr_pyramid_resident|corr_prep16" -s 4 -c 2 -o gpurun_out/prof_res -f python tools/tc_experiment.py child > gpurun_out/ncu_res.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_res.log
