#!/bin/bash
# round 2, call AC: ncu launch lists on the final tree (C5 line at 2048 lattices; C3 line without the also-legs)
set -x
mkdir -p gpurun_out
C5="python bench.py --workload c5 --envs 2048 --steps 100 --warmup 3 --no-cpu"
timeout 300 $C5 > gpurun_out/plain_c5.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c5.csv $C5 > gpurun_out/ncu_c5.log 2>&1
C3="python bench.py --steps 8 --warmup 3 --no-cpu --no-also --obs-to-host-steps 0"
timeout 300 $C3 > gpurun_out/plain_c3.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c3.csv $C3 > gpurun_out/ncu_c3.log 2>&1
ls -la gpurun_out/*.csv
