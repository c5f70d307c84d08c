#!/bin/bash
# round 2, call L: the driver's own command lines once, and the c2 per-call breakdown
set -x
mkdir -p gpurun_out
( time python3 bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/driver_ref.json 2> gpurun_out/driver_ref.err; echo "ref rc=$?"
( time python3 bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/driver_n1.json 2> gpurun_out/driver_n1.err; echo "n1 rc=$?"
tail -4 gpurun_out/driver_n1.err
python - <<PY
import json
for f in ["driver_ref", "driver_n1"]:
    d=json.loads([l for l in open("gpurun_out/%s.json"%f).read().strip().splitlines() if l.startswith("{")][-1])
    print(f, "%.4g"%d["value"], d.get("region_ms"), "e2e %.4g"%d["e2e"]["value"], "config", d["config"])
    for k,v in d.get("also",{}).items(): print("   also", k, v.get("value"), v.get("roofline",{}).get("frac"), v.get("e2e",{}).get("value"))
PY
timeout 300 python profiles/c2_breakdown.py cuda 3000 > gpurun_out/c2_breakdown_cuda.txt 2>&1; cat gpurun_out/c2_breakdown_cuda.txt
timeout 300 python profiles/c2_breakdown.py reference 3000 > gpurun_out/c2_breakdown_ref.txt 2>&1; cat gpurun_out/c2_breakdown_ref.txt
