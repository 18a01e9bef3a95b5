#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=/root/repo:/root/repo/tests
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x -k "forward or clip" 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
