#!/bin/bash
# usage: gpu_bench_n.sh N
N=${1:-2}
mkdir -p gpurun_out
export PYTHONPATH=/root/repo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"; tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'])
t=d['train']; print({k:t[k] for k in ('value','ms_per_step','global_batch','breakdown_ms')}, t['e2e'])
PY
