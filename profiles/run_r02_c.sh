#!/bin/bash
# round 2, call C: observation record (k_obs<CACHED>) + Ising environment interface
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
MFMARL_OBS_CACHED=1 timeout 900 python -m pytest tests/test_cuda_battle_abi.py tests/test_cuda_battle_batched.py tests/test_render_trace.py tests/test_algo_gpu.py -m gpu -x -q > gpurun_out/pytest_cached.log 2>&1; echo "pytest cached rc=$?" >> gpurun_out/pytest_cached.log
tail -6 gpurun_out/pytest_cached.log
timeout 300 python bench.py --workload c4 --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 rc=$?"
timeout 300 python bench.py --workload c4 --pipeline 1 --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4_p1.json 2> gpurun_out/bench_c4_p1.err; echo "c4 p1 rc=$?"
for t in 16 64 128; do timeout 300 python bench.py --workload c4 --obs-tile $t --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4_tile$t.json 2> gpurun_out/bench_c4_tile$t.err; done
MFMARL_OBS_CACHED=0 timeout 300 python bench.py --workload c4 --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4_uncached.json 2> gpurun_out/bench_c4_uncached.err
timeout 300 python bench.py --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "c3 rc=$?"
MFMARL_OBS_CACHED=1 timeout 300 python bench.py --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c3_cached.json 2> gpurun_out/bench_c3_cached.err; echo "c3 cached rc=$?"
for f in c4 c4_p1 c4_tile16 c4_tile64 c4_tile128 c4_uncached c3 c3_cached; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$f.json").read().strip().splitlines()[-1])
    print("$f", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], d.get("kernels_ms",{}).get("k_obs"), d.get("kernels_ms",{}).get("k_step"), d.get("kernels_alone_ms"), "frac", d["roofline"]["frac"] if "roofline" in d else None, "e2e %.4g"%d.get("e2e",{}).get("value"))
except Exception as ex:
    print("$f failed", ex)
PY
done
