#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_step_cl.csv python tools/profile_step.py step fp16 cl > gpurun_out/ncu_step_cl.log 2>&1; echo "rc=$?"
