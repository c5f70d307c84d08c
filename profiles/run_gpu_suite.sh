#!/bin/bash
# One gpurun call: GPU tests, the three bench lines, then the ncu passes (each only after its command exited 0 without ncu).
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "c3 rc=$?"
python bench.py --workload c5 --steps 100 --sweeps-per-launch 1 > gpurun_out/bench_c5_stream.json 2> gpurun_out/bench_c5_stream.err; echo "c5s rc=$?"
for S in 5 25 100; do
python bench.py --workload c5 --steps 100 --sweeps-per-launch $S --no-cpu > gpurun_out/bench_c5_res$S.json 2> gpurun_out/bench_c5_res$S.err; echo "c5r$S rc=$?"
done
python bench.py --workload c4 --no-cpu > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 rc=$?"
cat gpurun_out/bench_c3.json gpurun_out/bench_c5_stream.json gpurun_out/bench_c5_res*.json gpurun_out/bench_c4.json | cut -c1-600
# ncu: launch list of the default bench command, then one full capture of the resident Ising kernel
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c3.csv python bench.py --steps 8 --warmup 3 --no-cpu > gpurun_out/ncu_c3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_c5.csv python bench.py --workload c5 --envs 2048 --steps 50 --warmup 3 --no-cpu > gpurun_out/ncu_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ising_resident -s 2 -c 1 -o gpurun_out/ising_resident_r01 -f python bench.py --workload c5 --envs 2048 --steps 50 --warmup 3 --no-cpu > gpurun_out/ncu_c5_full.log 2>&1
ls -la gpurun_out
