#!/bin/bash
# round 2, call H: Ising persistent kernel v3 (early halo prefetch), full GPU suite, then the ncu evidence for profiles/r02
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --workload c5 --no-cpu > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "c5 rc=$?"
timeout 300 python bench.py --workload c5 --sweeps-per-launch 200 --no-cpu > gpurun_out/bench_c5_s200.json 2> gpurun_out/bench_c5_s200.err
for f in c5 c5_s200; do python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$f.json").read().strip().splitlines()[-1])
print("$f", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], d.get("region_ms"), "frac", d["roofline"]["frac"], "e2e %.4g"%d["e2e"]["value"])
PY
done
# ---- ncu (one tool per call): launch list of the bench command, then one full capture per dominant kernel
C3="python bench.py --steps 8 --warmup 3 --no-cpu --no-also --obs-to-host-steps 0"
timeout 300 $C3 > gpurun_out/plain_c3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_c3.csv $C3 > gpurun_out/ncu_launches_c3.log 2>&1
C4="python bench.py --workload c4 --pipeline 1 --obs-tile 32 --steps 8 --warmup 3 --no-cpu --obs-to-host-steps 0"
timeout 300 $C4 > gpurun_out/plain_c4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_obs -s 6 -c 1 -o gpurun_out/k_obs_c4_record -f $C4 > gpurun_out/ncu_k_obs_c4.log 2>&1
C5="python bench.py --workload c5 --envs 2048 --steps 100 --warmup 3 --sweeps-per-launch 25 --no-cpu"
timeout 300 $C5 > gpurun_out/plain_c5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_ising_persist -s 2 -c 1 -o gpurun_out/k_ising_persist -f $C5 > gpurun_out/ncu_k_ising_persist.log 2>&1
ls -la gpurun_out/*.ncu-rep
