#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest corr"; timeout 900 python -m pytest tests/test_gpu_corr.py tests/test_gpu_raft.py -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest.log
echo "=== experiment"; timeout 600 python tools/tc_experiment.py
