#!/bin/bash
# round 2, call M: C4 with fewer k_obs CTAs (room for the other engine's k_step on every SM?)
set -x
mkdir -p gpurun_out
for g in 0 232 264 280; do
MFMARL_OBS_GRID=$g timeout 300 python bench.py --workload c4 --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4_g$g.json 2> gpurun_out/bench_c4_g$g.err
MFMARL_OBS_GRID=$g timeout 300 python bench.py --workload c4 --obs-tile 32 --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4_t32_g$g.json 2> gpurun_out/bench_c4_t32_g$g.err
done
for f in c4_g0 c4_g232 c4_g264 c4_g280 c4_t32_g0 c4_t32_g232 c4_t32_g264 c4_t32_g280; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$f.json").read().strip().splitlines()[-1])
    print("$f", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], d.get("kernels_alone_ms"), "frac", d["roofline"]["frac"], "e2e %.4g"%d["e2e"]["value"])
except Exception as ex:
    print("$f failed", ex)
PY
done
