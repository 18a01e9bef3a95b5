#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=/root/repo:/root/repo/tests
timeout 900 python -m pytest $@ -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -25 | tee gpurun_out/pytest_one.log
