#!/bin/bash
# round 2, call AF: where the single-env k_step spends its 19 us -- source counters + warp-state sampling summed over
# 60 consecutive launches (step / clear_dead alternate), summarised on the box (the report itself is too large to keep)
set -x
mkdir -p gpurun_out
C2="python bench.py --workload c2 --steps 300 --warmup 20 --no-cpu"
timeout 900 ncu --section SourceCounters --section WarpStateStats --clock-control none --import-source on -k regex:k_step -s 100 -c 60 -o /tmp/k_step_c2_many -f $C2 > gpurun_out/ncu_k_step_c2_many.log 2>&1
ncu -i /tmp/k_step_c2_many.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src_many.csv 2>/dev/null
python profiles/src_lines.py /tmp/src_many.csv 1.0 > gpurun_out/k_step_c2_lines.txt
head -60 gpurun_out/k_step_c2_lines.txt
