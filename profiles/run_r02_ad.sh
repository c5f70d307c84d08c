#!/bin/bash
# round 2, call AD: the single-env loop under ncu -- launch list, then --set full of two consecutive k_step launches
# (the step launch and the speculative clear_dead launch) and one k_obs
set -x
mkdir -p gpurun_out
C2="python bench.py --workload c2 --steps 300 --warmup 20 --no-cpu"
timeout 300 $C2 > gpurun_out/plain_c2.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c2.csv $C2 > gpurun_out/ncu_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 100 -c 2 -o gpurun_out/k_step_c2 -f $C2 > gpurun_out/ncu_k_step_c2.log 2>&1
ls -la gpurun_out/*.ncu-rep
