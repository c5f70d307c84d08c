#!/bin/bash
# round 2, call T: the record lines on the final tree
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python3 bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/driver_ref.json 2> gpurun_out/driver_ref.err; echo "ref rc=$?"
python3 bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/driver_n1.json 2> gpurun_out/driver_n1.err; echo "n1 rc=$?"
python3 bench.py --workload c4 --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
python3 bench.py --workload c4 --pipeline 1 --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4_p1.json 2> gpurun_out/bench_c4_p1.err
python3 bench.py --workload c5 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
python3 bench.py --workload c5 --sweeps-per-launch 1 --steps 100 --no-cpu > gpurun_out/bench_c5_stream.json 2> gpurun_out/bench_c5_stream.err
python3 bench.py --pipeline 1 --no-cpu --no-also --obs-to-host-steps 0 > gpurun_out/bench_c3_p1.json 2> gpurun_out/bench_c3_p1.err
python - <<PY
import json
for f in ["driver_ref", "driver_n1", "bench_c4", "bench_c4_p1", "bench_c5", "bench_c5_stream", "bench_c3_p1"]:
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f).read().strip().splitlines() if l.startswith("{")][-1])
        print(f, "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], "frac", d.get("roofline",{}).get("frac"), "alone", (d.get("roofline",{}).get("alone") or {}).get("frac"), "e2e %.4g"%d["e2e"]["value"], (d.get("cpu_baseline") or {}).get("value"), (d.get("fresh_episodes") or {}).get("value"))
        for k,v in d.get("also",{}).items(): print("   also", k, v.get("value"), v.get("roofline",{}).get("frac"), v.get("e2e",{}).get("value"))
    except Exception as ex:
        print(f, "failed", ex)
PY
