#!/bin/bash
mkdir -p gpurun_out
for dbg in 15 7 3 2 0; do
  SDOF_TC_DEBUG=$dbg timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"corr_volume_tc|round_tf32|to_bf16" -s 12 -c 6 --csv --log-file gpurun_out/dur_$dbg.csv python tools/tc_experiment.py child > /dev/null 2>&1
  echo "== debug $dbg"; python - <<PY
import csv
rows=[l for l in open('gpurun_out/dur_$dbg.csv') if not l.startswith('==')]
out={}
for r in csv.DictReader(rows):
    out.setdefault((r['ID'],r['Kernel Name'][:40]),{})[r['Metric Name'][:40]]=r['Metric Value']
for k,v in out.items(): print(k, v)
PY
done
