#!/bin/bash
# GPU-box script: ncu captures behind the numbers in profiles/ (run under gpurun, one GPU).
#   usage: scripts/ncu_capture.sh <tag> [cfg ...]
# per config: (1) the plain run, (2) `--set full` tables of the step's kernels, (3) whole-step DRAM bytes with
# --cache-control none (the caches keep what the previous kernel left: the history is L2-resident when it fits).
tag=$1; shift
mkdir -p gpurun_out
for cfg in "$@"; do
  python scripts/one_step.py $cfg 4 > gpurun_out/plain_${tag}_$cfg.log 2>&1 || { echo "plain run of $cfg failed"; cat gpurun_out/plain_${tag}_$cfg.log; exit 1; }
  ncu --set full --clock-control none --import-source on -k regex:'k_walk|k_grad|k_emit' -s 6 -c 3 -o gpurun_out/prof_${tag}_$cfg -f python scripts/one_step.py $cfg 4 > gpurun_out/ncu_${tag}_$cfg.log 2>&1
  ncu -i gpurun_out/prof_${tag}_$cfg.ncu-rep --page raw --csv > gpurun_out/raw_${tag}_$cfg.csv 2>/dev/null
  ncu -i gpurun_out/prof_${tag}_$cfg.ncu-rep --page source --csv > gpurun_out/source_${tag}_$cfg.csv 2>/dev/null
  rm -f gpurun_out/prof_${tag}_$cfg.ncu-rep        # 55 MB each: only the exported tables travel back (gpurun_out is capped at 64 MiB)
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:'k_walk|k_grad|k_emit' -s 6 -c 3 --csv --log-file gpurun_out/traffic_${tag}_$cfg.csv python scripts/one_step.py $cfg 4 > /dev/null 2>&1
  echo "== $cfg"; tail -2 gpurun_out/ncu_${tag}_$cfg.log; grep -c k_ gpurun_out/traffic_${tag}_$cfg.csv
done
