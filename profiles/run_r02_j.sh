#!/bin/bash
# round 2, call J: record load by cp.async in k_obs<CACHED>, persistent halo buffer, full GPU suite
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
MFMARL_OBS_CACHED=1 timeout 900 python -m pytest tests/test_cuda_battle_abi.py tests/test_cuda_battle_batched.py -m gpu -x -q > gpurun_out/pytest_cached.log 2>&1; echo "pytest cached rc=$?" >> gpurun_out/pytest_cached.log
tail -3 gpurun_out/pytest_cached.log
timeout 300 python bench.py --workload c5 --no-cpu > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "c5 rc=$?"
timeout 300 python bench.py --workload c4 --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
for t in 32 64 96; do timeout 300 python bench.py --workload c4 --obs-tile $t --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4_t$t.json 2> gpurun_out/bench_c4_t$t.err; done
timeout 300 python bench.py --workload c4 --pipeline 1 --obs-tile 32 --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4_p1_t32.json 2> gpurun_out/bench_c4_p1_t32.err
for f in c5 c4 c4_t32 c4_t64 c4_t96 c4_p1_t32; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$f.json").read().strip().splitlines()[-1])
    print("$f", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], d.get("region_ms"), d.get("kernels_alone_ms"), "frac", d["roofline"]["frac"], "e2e %.4g"%d["e2e"]["value"], d["e2e"].get("region_ms"))
except Exception as ex:
    print("$f failed", ex)
PY
done
