#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=/root/repo:/root/repo/tests
timeout 300 python tools/f16_taps.py 2>&1 | tail -4 | tee gpurun_out/f16_taps.log
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -s -k "events or clip or f16 or dropout or host_path or train_then or non_finite" 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
