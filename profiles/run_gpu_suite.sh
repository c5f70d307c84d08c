#!/bin/bash
# One gpurun call: GPU tests, smoke, the bench lines, then the ncu passes (each only after its command exited 0
# without ncu).  Outputs land in gpurun_out/; the summaries kept for the record are copied to profiles/rNN/ by hand
# (profiles/summarize.py launches|raw, profiles/src_lines.py).
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "c3 rc=$?"
python bench.py --pipeline 1 --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c3_one_stream.json 2> gpurun_out/bench_c3_one_stream.err; echo "c3 p1 rc=$?"
python bench.py --workload c4 --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 rc=$?"
python bench.py --workload c5 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "c5 rc=$?"
python bench.py --workload c5 --sweeps-per-launch 1 --steps 100 --no-cpu > gpurun_out/bench_c5_stream.json 2> gpurun_out/bench_c5_stream.err; echo "c5 stream rc=$?"
python bench.py --workload c2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "c2 rc=$?"
for f in reference c3 c3_one_stream c4 c5 c5_stream c2; do cut -c1-300 gpurun_out/bench_$f.json; done
# ncu: launch list of the default bench command and of the Ising line, then one full capture per dominant kernel
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c3.csv python bench.py --steps 8 --warmup 3 --no-cpu --obs-to-host-steps 0 > gpurun_out/ncu_c3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_c5.csv python bench.py --workload c5 --envs 2048 --steps 100 --warmup 3 --no-cpu > gpurun_out/ncu_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_obs -s 6 -c 1 -o gpurun_out/k_obs -f python bench.py --steps 8 --warmup 3 --no-cpu --obs-to-host-steps 0 --pipeline 1 > gpurun_out/ncu_k_obs.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ising_resident -s 2 -c 1 -o gpurun_out/k_ising_resident -f python bench.py --workload c5 --envs 2048 --steps 100 --warmup 3 --sweeps-per-launch 25 --no-cpu > gpurun_out/ncu_k_ising.log 2>&1
ls -la gpurun_out
