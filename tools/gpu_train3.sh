#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=/root/repo:/root/repo/tests
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_train.log
for b in 64 128 256; do timeout 600 python tools/profile_train.py $b > gpurun_out/train_steps$b.txt 2>&1; head -1 gpurun_out/train_steps$b.txt; done
