#!/bin/bash
# round 2, call F: per-env placement + random sides, full GPU suite, default bench line
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "default rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
print("default", "%.4g"%d["value"], d["region_ms"], "frac", d["roofline"]["frac"], "e2e %.4g"%d["e2e"]["value"], d.get("cpu_baseline",{}).get("value"))
for k,v in d.get("also",{}).items(): print("   also", k, v.get("value"), v.get("roofline",{}).get("frac"), v.get("e2e",{}).get("value"))
PY
