#!/bin/bash
# ncu evidence: (1) launch list with device time per launch, (2) full capture of the tcgen05 GEMM kernel.
mkdir -p gpurun_out
export PYTHONPATH=/root/repo
python tools/ncu_target.py > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 236 -c 236 --csv --log-file gpurun_out/launches.csv \
    python tools/ncu_target.py > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
python tools/ncu_target.py > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 170 -c 4 -o gpurun_out/prof_gemm \
    python tools/ncu_target.py > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
tail -3 gpurun_out/ncu_plain.log gpurun_out/ncu_list.log gpurun_out/ncu_full.log
ls -la gpurun_out
