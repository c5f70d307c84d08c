#!/bin/bash
# round 2, call O: C4 tile sweep with the slot reservation, e2e included
set -x
mkdir -p gpurun_out
for t in 32 48 64 96; do timeout 300 python bench.py --workload c4 --obs-tile $t --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4_t$t.json 2> gpurun_out/bench_c4_t$t.err; done
for f in c4_t32 c4_t48 c4_t64 c4_t96; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$f.json").read().strip().splitlines()[-1])
    print("$f", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], d.get("kernels_alone_ms"), "frac", d["roofline"]["frac"], "e2e %.4g"%d["e2e"]["value"])
except Exception as ex:
    print("$f failed", ex)
PY
done
