#!/bin/bash
# Evidence for profiles/ (training plans): ncu launch list of one forward-with-tape + backward of 64 windows (second step of
# tools/ncu_target_train.py, plain stream) and ncu --set full of the two tensor-core backward kernels, summarised on the box.
mkdir -p gpurun_out
export PYTHONPATH=/root/repo
timeout 300 python tools/ncu_target_train.py > gpurun_out/ncu_train_plain.log 2>&1 || { tail -5 gpurun_out/ncu_train_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 660 -c 660 --csv --log-file gpurun_out/launches_train.csv \
    python tools/ncu_target_train.py > gpurun_out/ncu_train_list.log 2>&1
echo "list rc=$?"; wc -l gpurun_out/launches_train.csv
: > gpurun_out/full_metrics_train.txt; : > gpurun_out/stalls_by_line_train.txt
for spec in attn_local_bwd_tc_kernel:attn_local_bwd_tc:8 attn_global_bwd_kernel:attn_global_bwd:8 dwconv_ln_bwd_kernel:dwconv_ln_bwd_kernelILi128:30; do
  k=${spec%%:*}; rest=${spec#*:}; sec=${rest%%:*}; skip=${rest##*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$k" -s $skip -c 1 -o gpurun_out/fullt_$k -f \
      python tools/ncu_target_train.py > gpurun_out/ncu_fullt_$k.log 2>&1
  echo "$k rc=$?"
  echo "== $k" >> gpurun_out/full_metrics_train.txt; python tools/ncu_report.py gpurun_out/fullt_$k.ncu-rep 0 2>/dev/null | head -20 >> gpurun_out/full_metrics_train.txt
  echo "== $k" >> gpurun_out/stalls_by_line_train.txt; python tools/ncu_lines.py gpurun_out/fullt_$k.ncu-rep ${k%_kernel} 14 $sec 2>/dev/null >> gpurun_out/stalls_by_line_train.txt
  rm -f gpurun_out/fullt_$k.ncu-rep gpurun_out/ncu_fullt_$k.log
done
du -sh gpurun_out
