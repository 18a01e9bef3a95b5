#!/bin/bash
# Round-2 check: every -m gpu test, smoke(), full bench (our arm + reference arm).  Output under gpurun_out/.
mkdir -p gpurun_out
export PYTHONPATH=/root/repo:/root/repo/tests
set -o pipefail
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader | head -2
nproc; numactl -H 2>/dev/null | head -4; cat /sys/devices/system/node/online 2>/dev/null
timeout 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider -s 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
