#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=/root/repo:/root/repo/tests
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'])
print(d['roofline']['families_ms'], 'frac', d['roofline']['frac'], d['clocks'])
print(json.dumps(d.get('train'), indent=1))
PY
tail -3 gpurun_out/bench.err
