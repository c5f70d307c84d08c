#!/bin/bash
# round 2, call W: Ising persistent kernel K6s (SWAR neighbour counts) -- tests, bench, old kernel beside it
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cuda_ising.py tests/test_ising_env.py -m gpu -x -q > gpurun_out/pytest_ising.log 2>&1; echo "pytest ising rc=$?" >> gpurun_out/pytest_ising.log
tail -5 gpurun_out/pytest_ising.log
timeout 300 python bench.py --workload c5 --no-cpu > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "c5 rc=$?"
MFMARL_ISING_PERSIST=1 timeout 300 python bench.py --workload c5 --no-cpu > gpurun_out/bench_c5_k6p.json 2> gpurun_out/bench_c5_k6p.err; echo "c5 k6p rc=$?"
for f in c5 c5_k6p; do python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$f.json").read().strip().splitlines()[-1])
print("$f", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], d.get("region_ms"), "frac", d["roofline"]["frac"], "e2e %.4g"%d["e2e"]["value"])
PY
done
