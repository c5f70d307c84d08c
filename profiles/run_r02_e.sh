#!/bin/bash
# round 2, call E: single-env ABI with speculation (tests + c2 latency), C4 pipeline/tile variants, c5 repeats
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --workload c2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "c2 rc=$?"
MAGENT_SPECULATE=0 timeout 300 python bench.py --workload c2 --no-cpu > gpurun_out/bench_c2_nospec.json 2> gpurun_out/bench_c2_nospec.err; echo "c2 nospec rc=$?"
timeout 300 python bench.py --workload c4 --pipeline 4 --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4_p4.json 2> gpurun_out/bench_c4_p4.err
timeout 300 python bench.py --workload c4 --pipeline 4 --obs-tile 64 --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4_p4_t64.json 2> gpurun_out/bench_c4_p4_t64.err
timeout 300 python bench.py --workload c4 --obs-tile 256 --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4_t256.json 2> gpurun_out/bench_c4_t256.err
timeout 300 python bench.py --workload c4 --obs-tile 96 --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4_t96.json 2> gpurun_out/bench_c4_t96.err
timeout 300 python bench.py --workload c4 --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
timeout 300 python bench.py --workload c5 --no-cpu > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "c5 rc=$?"
for f in c2 c2_nospec c4_p4 c4_p4_t64 c4_t256 c4_t96 c4 c5; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$f.json").read().strip().splitlines()[-1])
    print("$f", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], d.get("region_ms"), d.get("kernels_alone_ms"), "frac", d.get("roofline",{}).get("frac"), "e2e %.4g"%d.get("e2e",{}).get("value"), d.get("cpu_baseline",{}).get("value"))
except Exception as ex:
    print("$f failed", ex)
PY
done
