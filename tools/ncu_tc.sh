#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:corr_volume_tc -s 3 -c 1 -o gpurun_out/prof_tc -f python tools/tc_experiment.py child > gpurun_out/ncu_tc.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_tc.log; ls -la gpurun_out/prof_tc.ncu-rep
