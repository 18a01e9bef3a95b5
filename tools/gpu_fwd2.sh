#!/bin/bash
# forward parity tests + per-step profile (+ FFN timeline when the library was built with -DA2M_FFN_TIMING)
mkdir -p gpurun_out
export PYTHONPATH=/root/repo:/root/repo/tests
timeout 600 python -m pytest tests/test_gpu_forward.py tests/test_gpu_events.py tests/test_gpu_clip.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/pytest_fwd.log
timeout 300 python tools/profile_steps.py 64 > gpurun_out/steps64.txt 2>&1; head -1 gpurun_out/steps64.txt; sed -n 50,66p gpurun_out/steps64.txt
timeout 200 python tools/ffn_timeline.py 64 2>&1 | tail -12
