#!/bin/bash
# forward parity tests + per-step profile + short forward-only bench
mkdir -p gpurun_out
export PYTHONPATH=/root/repo:/root/repo/tests
timeout 600 python -m pytest tests/test_gpu_forward.py tests/test_gpu_events.py tests/test_gpu_clip.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/pytest_fwd.log
timeout 300 python tools/profile_steps.py 64 > gpurun_out/steps64.txt 2>&1; head -1 gpurun_out/steps64.txt; grep -E "ffn_fused|attn_|qkv" gpurun_out/steps64.txt | head -8
timeout 600 python bench.py --steps 20 --warmup 3 --no-train > gpurun_out/bench_fwd.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_fwd.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'])
print(d['roofline']['families_ms'], 'frac', d['roofline']['frac'], d['clocks'])
PY
tail -3 gpurun_out/bench.err
