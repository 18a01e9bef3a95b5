#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=/root/repo:/root/repo/tests
timeout 300 python tools/train_trace.py 64 8 0.1 2>&1 | cut -c1-260 > gpurun_out/train_trace.txt; head -12 gpurun_out/train_trace.txt
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_train.log
