#!/bin/bash
# parity tests + per-step profile + short bench
mkdir -p gpurun_out
export PYTHONPATH=/root/repo:/root/repo/tests
timeout 900 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 300 python tools/profile_steps.py 64 > gpurun_out/steps64.txt 2>&1; head -3 gpurun_out/steps64.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'])
print(d['roofline']['families_ms'], 'frac', d['roofline']['frac'], d['clocks'])
PY
tail -3 gpurun_out/bench.err
