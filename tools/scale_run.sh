#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
echo "=== N=1"; timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"; tail -2 gpurun_out/bench_n1.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, d['e2e'], d['clocks'], 'roofline', d['roofline']['frac'], d['roofline']['us_per_launch'])"
echo "=== N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"; tail -3 gpurun_out/bench_n$N.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_n$N.json').read().strip().splitlines() if l.startswith('{')][-1]); print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, d['e2e'])"
echo "=== reference arm under torchrun"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 10 --warmup 2 2>/dev/null | grep impl | cut -c1-200
