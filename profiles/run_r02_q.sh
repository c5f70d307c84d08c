#!/bin/bash
# round 2, call Q: what the driver runs at round end, on the final tree -- GPU tests, smoke, reference arm, default line
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
python3 bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/driver_ref.json 2> gpurun_out/driver_ref.err; echo "ref rc=$?"
python3 bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/driver_n1.json 2> gpurun_out/driver_n1.err; echo "n1 rc=$?"
python3 bench.py --workload c2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "c2 rc=$?"
python - <<PY
import json
for f in ["driver_ref", "driver_n1", "bench_c2"]:
    d=json.loads([l for l in open("gpurun_out/%s.json"%f).read().strip().splitlines() if l.startswith("{")][-1])
    print(f, "%.4g"%d["value"], d.get("region_ms"), "frac", d.get("roofline",{}).get("frac"), "e2e %.4g"%d["e2e"]["value"], (d.get("cpu_baseline") or {}).get("value"))
    for k,v in d.get("also",{}).items(): print("   also", k, v.get("value"), v.get("roofline",{}).get("frac"), v.get("e2e",{}).get("value"))
PY
