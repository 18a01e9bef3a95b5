#!/bin/bash
# First-light GPU check: every stage in its own process under a timeout so a fault in one does not hide the rest.
mkdir -p gpurun_out
export PYTHONPATH=/root/repo:/root/repo/tests
nvidia-smi --query-gpu=name,memory.total --format=csv | tee gpurun_out/gpu.txt
for t in tests/test_gpu_gemm.py tests/test_gpu_forward.py; do
  echo "=== $t"
  timeout 600 python -m pytest $t -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -25
done 2>&1 | tee gpurun_out/first_light.log
