#!/bin/bash
# round 2, call D: Ising persistent kernel v2 (boundary row first, prefetched halo), the new bench line (repeated
# regions + also-legs), k_step on high-priority streams (experiment)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cuda_ising.py tests/test_ising_env.py -m gpu -x -q > gpurun_out/pytest_ising.log 2>&1; echo "pytest ising rc=$?" >> gpurun_out/pytest_ising.log
tail -3 gpurun_out/pytest_ising.log
timeout 300 python bench.py --workload c5 --no-cpu > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "c5 rc=$?"
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "default rc=$?"
timeout 300 python bench.py --workload c4 --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 rc=$?"
timeout 300 python bench.py --workload c4 --step-priority --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4_prio128.json 2> gpurun_out/bench_c4_prio128.err
timeout 300 python bench.py --workload c4 --step-priority --obs-tile 32 --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4_prio32.json 2> gpurun_out/bench_c4_prio32.err
timeout 300 python bench.py --workload c4 --step-priority --obs-tile 64 --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4_prio64.json 2> gpurun_out/bench_c4_prio64.err
timeout 300 python bench.py --workload c4 --pipeline 1 --obs-tile 128 --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4_p1_128.json 2> gpurun_out/bench_c4_p1_128.err
timeout 300 python bench.py --step-priority --no-cpu --no-also --obs-to-host-steps 0 > gpurun_out/bench_c3_prio.json 2> gpurun_out/bench_c3_prio.err
for f in c5 default c4 c4_prio128 c4_prio32 c4_prio64 c4_p1_128 c3_prio; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$f.json").read().strip().splitlines()[-1])
    print("$f", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], d.get("region_ms"), d.get("kernels_alone_ms"), "frac", d["roofline"]["frac"], "e2e %.4g"%d.get("e2e",{}).get("value"))
    for k,v in d.get("also",{}).items(): print("   also", k, v.get("value"), v.get("roofline",{}).get("frac"), v.get("e2e",{}).get("value"))
except Exception as ex:
    print("$f failed", ex)
PY
done
tail -3 gpurun_out/bench_default.err
