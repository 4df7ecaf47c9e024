#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_m.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_m.log; grep "720x1280" gpurun_out/pytest_m.log
echo "=== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1]); print('N=2 value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'clip',d['clip']['value'],'batched',d['batched']['value'])"; tail -3 gpurun_out/bench_n2.err
echo "=== ref arm N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 2>/dev/null | cut -c1-160
