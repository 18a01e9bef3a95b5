#!/bin/bash
# Evidence for profiles/: (1) ncu launch list of the bench command, (2) ncu --set full of the top kernels, summarised ON
# THE BOX (the .ncu-rep files together exceed the 64 MiB that gpurun copies back; only the first one is kept)
mkdir -p gpurun_out
export PYTHONPATH=/root/repo
timeout 600 python bench.py --steps 2 --warmup 1 --no-train > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-train > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"; wc -l gpurun_out/launches.csv
python tools/ncu_target.py > gpurun_out/ncu_plain.log 2>&1 || exit 1
: > gpurun_out/full_metrics.txt; : > gpurun_out/stalls_by_line.txt
for spec in block_fused_kernel:block_fused_kernelILi128ELb0 postattn_fused_kernel:postattn_fused qkv_fused_kernel:qkv_fused attn_global_kernel:attn_global \
            block_mid_kernel:block_mid_kernelILi32 attn_local_tc_kernel:attn_local_tc block256_fused_kernel:block256_fused; do
  k=${spec%%:*}; sec=${spec##*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$k" -s 4 -c 1 -o gpurun_out/full_$k -f \
      python tools/ncu_target.py > gpurun_out/ncu_full_$k.log 2>&1
  echo "$k rc=$?"
  echo "== $k" >> gpurun_out/full_metrics.txt; python tools/ncu_report.py gpurun_out/full_$k.ncu-rep 0 2>/dev/null | head -20 >> gpurun_out/full_metrics.txt
  echo "== $k" >> gpurun_out/stalls_by_line.txt; python tools/ncu_lines.py gpurun_out/full_$k.ncu-rep ${k%_kernel} 14 $sec 2>/dev/null >> gpurun_out/stalls_by_line.txt
  [ "$k" != block_fused_kernel ] && rm -f gpurun_out/full_$k.ncu-rep
done
rm -f gpurun_out/ncu_full_*.log
du -sh gpurun_out
