#!/bin/bash
# round 2, call B: persistent Ising kernel (first run), k_step with the 32-bit grid, ncu source profile of k_step at C3
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cuda_ising.py -m gpu -x -q > gpurun_out/pytest_ising.log 2>&1; echo "pytest ising rc=$?" >> gpurun_out/pytest_ising.log
tail -5 gpurun_out/pytest_ising.log
timeout 600 python -m pytest tests/test_cuda_battle_abi.py tests/test_cuda_battle_batched.py -m gpu -x -q > gpurun_out/pytest_battle.log 2>&1; echo "pytest battle rc=$?" >> gpurun_out/pytest_battle.log
tail -5 gpurun_out/pytest_battle.log
timeout 300 python bench.py --workload c5 --no-cpu > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "c5 rc=$?"
MFMARL_ISING_PERSIST=0 timeout 300 python bench.py --workload c5 --no-cpu > gpurun_out/bench_c5_cluster.json 2> gpurun_out/bench_c5_cluster.err; echo "c5 cluster rc=$?"
timeout 300 python bench.py --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "c3 rc=$?"
timeout 300 python bench.py --pipeline 1 --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c3_p1.json 2> gpurun_out/bench_c3_p1.err; echo "c3 p1 rc=$?"
timeout 300 python bench.py --step-threads 128 --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c3_t128.json 2> gpurun_out/bench_c3_t128.err
timeout 300 python bench.py --workload c4 --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 rc=$?"
timeout 300 python bench.py --workload c4 --step-threads 1024 --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4_t1024.json 2> gpurun_out/bench_c4_t1024.err
timeout 300 python bench.py --workload c4 --step-threads 256 --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4_t256.json 2> gpurun_out/bench_c4_t256.err
for f in c5 c5_cluster c3 c3_p1 c3_t128 c4 c4_t1024 c4_t256; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$f.json").read().strip().splitlines()[-1])
    print("$f", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], d.get("kernels_ms",{}).get("k_obs"), d.get("kernels_ms",{}).get("k_step"), d.get("kernels_alone_ms"), "frac", d["roofline"]["frac"] if "roofline" in d else None, "e2e %.4g"%d.get("e2e",{}).get("value"))
except Exception as ex:
    print("$f failed", ex)
PY
done
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu --obs-to-host-steps 0 --pipeline 1 > gpurun_out/plain_c3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_step -s 6 -c 1 -o gpurun_out/k_step_c3 -f python bench.py --steps 8 --warmup 3 --no-cpu --obs-to-host-steps 0 --pipeline 1 > gpurun_out/ncu_k_step_c3.log 2>&1
timeout 300 python bench.py --workload c4 --steps 8 --warmup 3 --no-cpu --obs-to-host-steps 0 --pipeline 1 > gpurun_out/plain_c4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_step -s 6 -c 1 -o gpurun_out/k_step_c4 -f python bench.py --workload c4 --steps 8 --warmup 3 --no-cpu --obs-to-host-steps 0 --pipeline 1 > gpurun_out/ncu_k_step_c4.log 2>&1
ls -la gpurun_out
