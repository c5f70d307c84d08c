#!/bin/bash
# round 2, call S: fresh-episode sub-measurement, bf16 rollout in training test, obs rows probe
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_algo_gpu.py -m gpu -x -q > gpurun_out/pytest_algo.log 2>&1; echo "pytest algo rc=$?" >> gpurun_out/pytest_algo.log
tail -4 gpurun_out/pytest_algo.log
python3 bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/driver_n1.json 2> gpurun_out/driver_n1.err; echo "n1 rc=$?"
tail -3 gpurun_out/driver_n1.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/driver_n1.json").read().strip().splitlines() if l.startswith("{")][-1])
print("n1", "%.4g"%d["value"], d["agents_per_step"], d.get("region_ms"), "frac", d["roofline"]["frac"], "e2e %.4g"%d["e2e"]["value"])
print("fresh", d["fresh_episodes"])
for k,v in d.get("also",{}).items(): print("   also", k, v.get("value"), v.get("roofline",{}).get("frac"), v.get("e2e",{}).get("value"))
PY
timeout 300 python profiles/obs_rows_probe.py > gpurun_out/obs_rows_probe.txt 2>&1; cat gpurun_out/obs_rows_probe.txt
