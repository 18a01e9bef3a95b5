#!/bin/bash
# usage: gpu_ncu_train.sh "<kernel regex>" <skip> <count> <outname>
mkdir -p gpurun_out
export PYTHONPATH=/root/repo
K="$1"; S="${2:-0}"; C="${3:-1}"; O="${4:-prof_t}"
python tools/ncu_train_target.py > gpurun_out/ncu_plain.log 2>&1 || { tail -5 gpurun_out/ncu_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"$K" -s $S -c $C -o gpurun_out/$O -f \
    python tools/ncu_train_target.py > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"; tail -n 2 gpurun_out/ncu_full.log
