#!/bin/bash
# round 2, call AE: final tree -- GPU tests, smoke, the driver's two command lines, C2, the rollout lines
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
python3 bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/driver_ref.json 2> gpurun_out/driver_ref.err; echo "ref rc=$?"
python3 bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/driver_n1.json 2> gpurun_out/driver_n1.err; echo "n1 rc=$?"
python3 bench.py --workload c2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "c2 rc=$?"
for a in mfq mfac; do python3 bench.py --workload play --algo $a --policy-precision bf16rows > gpurun_out/bench_play_${a}_bf16rows.json 2> /dev/null; done
timeout 300 python profiles/play_busy_probe.py 2>&1 | grep -v -i warn > gpurun_out/play_busy_probe.txt
python - <<PY
import json
for f in ["driver_ref", "driver_n1", "bench_c2", "bench_play_mfq_bf16rows", "bench_play_mfac_bf16rows"]:
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f).read().strip().splitlines() if l.startswith("{")][-1])
        print(f, "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], "frac", d.get("roofline",{}).get("frac"), "e2e", (d.get("e2e") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"))
        for k,v in d.get("also",{}).items(): print("   also", k, v.get("value"), v.get("roofline",{}).get("frac"), v.get("e2e",{}).get("value"))
    except Exception as ex:
        print(f, "failed", ex)
PY
