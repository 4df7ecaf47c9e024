#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_write.sum,l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"corr_pyramid_resident|corr_prep16" -s 4 -c 4 --csv --log-file gpurun_out/dur_res.csv python tools/tc_experiment.py child > /dev/null 2>&1
python - <<PY
import csv
rows=[l for l in open('gpurun_out/dur_res.csv') if not l.startswith('==')]
out={}
for r in csv.DictReader(rows):
    out.setdefault((r['ID'],r['Kernel Name'][:30]),{})[r['Metric Name'][:34]]=r['Metric Value']
for k,v in out.items(): print(k, v)
PY
