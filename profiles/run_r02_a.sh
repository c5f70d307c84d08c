#!/bin/bash
# round 2, call A: GPU tests with the rewritten k_step, then the bench lines that show what it bought
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
profiles/bin/cluster_probe > gpurun_out/cluster_probe.txt 2>&1; cat gpurun_out/cluster_probe.txt
timeout 300 python bench.py --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "c3 rc=$?"
timeout 300 python bench.py --pipeline 1 --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c3_p1.json 2> gpurun_out/bench_c3_p1.err; echo "c3 p1 rc=$?"
timeout 300 python bench.py --workload c4 --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 rc=$?"
timeout 300 python bench.py --workload c4 --pipeline 1 --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4_p1.json 2> gpurun_out/bench_c4_p1.err; echo "c4 p1 rc=$?"
for t in 32 64 256; do timeout 300 python bench.py --workload c4 --obs-tile $t --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4_tile$t.json 2> gpurun_out/bench_c4_tile$t.err; done
timeout 300 python bench.py --workload c2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "c2 rc=$?"
for f in c3 c3_p1 c4 c4_p1 c4_tile32 c4_tile64 c4_tile256 c2; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d.get("kernels_ms"), d.get("kernels_alone_ms"), d["roofline"]["frac"] if "roofline" in d else None, d.get("e2e",{}).get("value"))
except Exception as ex:
    print("$f failed", ex)
PY
done
