#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=/root/repo
python tools/ncu_target.py > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"block_fused_kernel|attn_local|attn_global" -s 32 -c 6 -o gpurun_out/prof_b \
    python tools/ncu_target.py > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"; tail -n 3 gpurun_out/ncu_plain.log; tail -n 3 gpurun_out/ncu_full.log
