#!/bin/bash
# round 2, call V: fresh --set full capture of the persistent Ising kernel (stall reasons per line)
set -x
mkdir -p gpurun_out
C5="python bench.py --workload c5 --envs 2048 --steps 100 --warmup 3 --sweeps-per-launch 50 --no-cpu"
timeout 300 $C5 > gpurun_out/plain_c5.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ising_persist -s 2 -c 1 -o gpurun_out/k_ising_swar_v2 -f $C5 > gpurun_out/ncu_k_ising_persist.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
