#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=/root/repo:/root/repo/tests
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -x --no-header -p no:cacheprovider -k wgrad 2>&1 | tail -15 | tee gpurun_out/pytest_wgrad.log
timeout 900 python tools/check_grads.py 2 > gpurun_out/check_grads.txt 2>&1; tail -5 gpurun_out/check_grads.txt
