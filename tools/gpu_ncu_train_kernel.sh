#!/bin/bash
# one `ncu --set full` capture of a training-plan kernel: bash tools/gpu_ncu_train_kernel.sh <kernel regex> <skip> <count>
mkdir -p gpurun_out
export PYTHONPATH=/root/repo
K="${1:-block_mid_bwd_kernel}"
S="${2:-0}"
N="${3:-1}"
timeout 300 python tools/ncu_target_train.py > gpurun_out/ncu_train_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s $S -c $N -o gpurun_out/prof_train_k -f \
    python tools/ncu_target_train.py > gpurun_out/ncu_train_full.log 2>&1
echo "full rc=$?"; tail -n 2 gpurun_out/ncu_train_plain.log; tail -n 2 gpurun_out/ncu_train_full.log
