#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=/root/repo:/root/repo/tests
timeout 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider -s "$@" 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
