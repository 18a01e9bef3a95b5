#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=/root/repo:/root/repo/tests
timeout 900 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
timeout 600 python tools/profile_train.py 64 all > gpurun_out/train_steps64.txt 2>&1; head -40 gpurun_out/train_steps64.txt
timeout 300 python tools/profile_train.py 8 > gpurun_out/train_steps8.txt 2>&1; head -3 gpurun_out/train_steps8.txt
