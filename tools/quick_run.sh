#!/bin/bash
# tests + diagnostics (no ncu)
mkdir -p gpurun_out
echo "=== pytest"; timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest.log
echo "=== diag"; timeout 900 python tools/gpu_diag.py > gpurun_out/diag_stdout.log 2>&1; echo "diag rc=$?"; grep -v "^(" gpurun_out/diag.txt | tail -40
