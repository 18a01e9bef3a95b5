#!/bin/bash
# Evidence for profiles/: (1) ncu launch list of the bench command, (2) ncu --set full of the top kernels
mkdir -p gpurun_out
export PYTHONPATH=/root/repo
timeout 600 python bench.py --steps 2 --warmup 1 --no-train > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-train > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"; wc -l gpurun_out/launches.csv
python tools/ncu_target.py > gpurun_out/ncu_plain.log 2>&1 || exit 1
for k in block_fused_kernel postattn_fused_kernel qkv_fused_kernel attn_global_kernel block_mid_kernel attn_local_tc_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$k" -s 4 -c 1 -o gpurun_out/full_$k -f \
      python tools/ncu_target.py > gpurun_out/ncu_full_$k.log 2>&1
  echo "$k rc=$?"
done
ls -la gpurun_out | tail -12
