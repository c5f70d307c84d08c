#!/bin/bash
# round 2, call X: K6s at 64 / 128 / 256 -- tests, C5 line, probes beside the previous resident kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cuda_ising.py tests/test_ising_env.py -m gpu -x -q > gpurun_out/pytest_ising.log 2>&1; echo "pytest ising rc=$?" >> gpurun_out/pytest_ising.log
tail -5 gpurun_out/pytest_ising.log
timeout 300 python bench.py --workload c5 --no-cpu > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "c5 rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_c5.json").read().strip().splitlines()[-1])
print("c5", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e %.4g"%d["e2e"]["value"])
PY
{
for L in 512 128 64; do
  B=$((16384 * 65536 / L / L / 8))
  echo "== L=$L K6s"; timeout 120 python profiles/ising_probe.py $B 100 $L resident
  echo "== L=$L previous resident kernel"; MFMARL_ISING_PERSIST=1 timeout 120 python profiles/ising_probe.py $B 100 $L resident
  echo "== L=$L streaming"; timeout 120 python profiles/ising_probe.py $B 20 $L streaming
done
} > gpurun_out/ising_probe_sizes.txt 2>&1
cat gpurun_out/ising_probe_sizes.txt
