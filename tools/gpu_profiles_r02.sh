#!/bin/bash
# Round-2 evidence for profiles/ (forward): (1) ncu launch list of the bench command, (2) ncu --set full of the largest kernels on
# tools/ncu_target.py (plain-stream forwards of 64 windows), summarised ON THE BOX (the reports together exceed what gpurun copies back).
mkdir -p gpurun_out
export PYTHONPATH=/root/repo
export A2M_PROFILE_LIB=/root/repo/audio-to-midi_b200/_build/libaudio2midi_b200_f16.so   # inference runs on the f16-operand build
BENCH="python bench.py --steps 2 --warmup 1 --no-train --no-extra --no-cpu"
timeout 600 $BENCH > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
    $BENCH > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"; wc -l gpurun_out/launches.csv
python tools/ncu_target.py > gpurun_out/ncu_plain.log 2>&1 || { tail -5 gpurun_out/ncu_plain.log; exit 1; }
: > gpurun_out/full_metrics.txt; : > gpurun_out/stalls_by_line.txt; : > gpurun_out/traffic.jsonl
for spec in "block_fused_kernel|block_fused_kernel<\(int\)128, \(bool\)0>|block_fused_kernelILi128ELb0" \
            "postattn_fused_kernel|postattn_fused_kernel|postattn_fused" "qkv_fused_kernel|qkv_fused_kernel|qkv_fused" \
            "attn_global_kernel|attn_global_kernel|attn_global" "attn_local_tc_kernel|attn_local_tc_kernel|attn_local_tc" \
            "block256_fused_kernel|block256_fused_kernel|block256_fused" "block_mid_kernel|block_mid2_kernel<\(int\)32>|block_mid2_kernelILi32" \
            "stage_small_kernel|stage_small_kernel<\(int\)4>|stage_small_kernelILi4"; do
  fam=${spec%%|*}; rest=${spec#*|}; rx=${rest%%|*}; sec=${rest##*|}
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$rx" -s 3 -c 1 -o gpurun_out/full_$fam -f \
      python tools/ncu_target.py > gpurun_out/ncu_full_$fam.log 2>&1
  echo "$fam rc=$?"
  echo "== $fam" >> gpurun_out/full_metrics.txt; python tools/ncu_report.py gpurun_out/full_$fam.ncu-rep 0 2>/dev/null | head -20 >> gpurun_out/full_metrics.txt
  echo "== $fam" >> gpurun_out/stalls_by_line.txt; python tools/ncu_lines.py gpurun_out/full_$fam.ncu-rep ${fam%_kernel} 14 $sec 2>/dev/null >> gpurun_out/stalls_by_line.txt
  python tools/ncu_traffic.py gpurun_out/full_$fam.ncu-rep $fam >> gpurun_out/traffic.jsonl 2>/dev/null
  [ "$fam" != block_fused_kernel ] && rm -f gpurun_out/full_$fam.ncu-rep
done
rm -f gpurun_out/ncu_full_*.log
cat gpurun_out/traffic.jsonl; head -30 gpurun_out/full_metrics.txt; du -sh gpurun_out
