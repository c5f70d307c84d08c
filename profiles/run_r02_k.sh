#!/bin/bash
# round 2, call K: Ising persistent kernel v5 (temperature table, REDUX statistics) + ncu capture of it
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cuda_ising.py tests/test_ising_env.py -m gpu -x -q > gpurun_out/pytest_ising.log 2>&1; echo "pytest ising rc=$?" >> gpurun_out/pytest_ising.log
tail -3 gpurun_out/pytest_ising.log
timeout 300 python bench.py --workload c5 --no-cpu > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "c5 rc=$?"
timeout 300 python bench.py --workload c5 --sweeps-per-launch 200 --no-cpu > gpurun_out/bench_c5_s200.json 2> gpurun_out/bench_c5_s200.err
for f in c5 c5_s200; do python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$f.json").read().strip().splitlines()[-1])
print("$f", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], d.get("region_ms"), "frac", d["roofline"]["frac"], "e2e %.4g"%d["e2e"]["value"])
PY
done
C5="python bench.py --workload c5 --envs 2048 --steps 100 --warmup 3 --sweeps-per-launch 50 --no-cpu"
timeout 300 $C5 > gpurun_out/plain_c5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_ising_persist -s 2 -c 1 -o gpurun_out/k_ising_persist_v5 -f $C5 > gpurun_out/ncu_k_ising_persist.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
