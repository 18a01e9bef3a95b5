#!/bin/bash
# Round-2 evidence for profiles/ (training plans): ncu launch list of one forward-with-tape + backward of 64 windows (second step of
# tools/ncu_target_train.py, plain stream) and ncu --set full of the weight-gradient GEMM, the dgrad GEMM and the backward attention /
# depthwise kernels, summarised on the box.
mkdir -p gpurun_out
export PYTHONPATH=/root/repo
timeout 300 python tools/ncu_target_train.py > gpurun_out/ncu_train_plain.log 2>&1 || { tail -5 gpurun_out/ncu_train_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 660 -c 660 --csv --log-file gpurun_out/launches_train.csv \
    python tools/ncu_target_train.py > gpurun_out/ncu_train_list.log 2>&1
echo "list rc=$?"; wc -l gpurun_out/launches_train.csv
: > gpurun_out/full_metrics_train.txt; : > gpurun_out/stalls_by_line_train.txt; : > gpurun_out/traffic_train.jsonl
for spec in "gemm_wgrad_kernel|gemm_wgrad_kernel<\(int\)256>|gemm_wgrad_kernelILi256|40" \
            "gemm_tc2_kernel|gemm_tc2_kernel<\(int\)256, \(int\)1, \(bool\)0>|gemm_tc2_kernelILi256ELi1ELb0|40" \
            "attn_local_bwd_tc_kernel|attn_local_bwd_tc_kernel|attn_local_bwd_tc|8" "attn_global_bwd_kernel|attn_global_bwd_kernel|attn_global_bwd|8" \
            "dwconv_ln_bwd_kernel|dwconv_ln_bwd_kernel<\(int\)128>|dwconv_ln_bwd_kernelILi128|25" \
            "block_fused_kernel|block_fused_kernel<\(int\)128, \(bool\)1>|block_fused_kernelILi128ELb1|25"; do
  fam=${spec%%|*}; rest=${spec#*|}; rx=${rest%%|*}; rest=${rest#*|}; sec=${rest%%|*}; skip=${rest##*|}
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$rx" -s $skip -c 1 -o gpurun_out/fullt_$fam -f \
      python tools/ncu_target_train.py > gpurun_out/ncu_fullt_$fam.log 2>&1
  echo "$fam rc=$?"
  echo "== $fam ($rx)" >> gpurun_out/full_metrics_train.txt; python tools/ncu_report.py gpurun_out/fullt_$fam.ncu-rep 0 2>/dev/null | head -20 >> gpurun_out/full_metrics_train.txt
  echo "== $fam ($rx)" >> gpurun_out/stalls_by_line_train.txt; python tools/ncu_lines.py gpurun_out/fullt_$fam.ncu-rep ${fam%_kernel} 14 $sec 2>/dev/null >> gpurun_out/stalls_by_line_train.txt
  python tools/ncu_traffic.py gpurun_out/fullt_$fam.ncu-rep $fam >> gpurun_out/traffic_train.jsonl 2>/dev/null
  rm -f gpurun_out/fullt_$fam.ncu-rep gpurun_out/ncu_fullt_$fam.log
done
cat gpurun_out/traffic_train.jsonl; head -45 gpurun_out/full_metrics_train.txt; du -sh gpurun_out
