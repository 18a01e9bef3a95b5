#!/bin/bash
# timeline build: forward + training parity tests, per-step profile, timelines
mkdir -p gpurun_out
export PYTHONPATH=/root/repo:/root/repo/tests
timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_events.py tests/test_gpu_clip.py tests/test_gpu_train.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/pytest_fwd.log
timeout 300 python tools/profile_steps.py 64 > gpurun_out/steps64.txt 2>&1; head -1 gpurun_out/steps64.txt; sed -n 56,64p gpurun_out/steps64.txt
timeout 200 python tools/ffn_timeline.py 64 2>&1 | tail -3
