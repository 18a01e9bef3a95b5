#!/bin/bash
# Full GPU session: training trace, parity tests, smoke, bench, reference arm.
mkdir -p gpurun_out
export PYTHONPATH=/root/repo:/root/repo/tests
set -o pipefail
timeout 300 python tools/train_trace.py 64 8 0.1 > gpurun_out/train_trace.txt 2>&1; tail -12 gpurun_out/train_trace.txt
timeout 300 python tools/train_trace.py 64 4 0.0 > gpurun_out/train_trace_nodrop.txt 2>&1; tail -6 gpurun_out/train_trace_nodrop.txt
timeout 900 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
