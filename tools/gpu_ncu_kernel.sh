#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=/root/repo
K="${1:-block_fused_kernel}"
S="${2:-10}"
python tools/ncu_target.py > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$K" -s $S -c 2 -o gpurun_out/prof_k \
    python tools/ncu_target.py > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"; tail -n 2 gpurun_out/ncu_full.log
