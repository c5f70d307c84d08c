#!/bin/bash
# round 2, call N: bf16 observation rows + bf16 rollout twins, slot reservation for pipelined engines, full GPU suite
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --workload c4 --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 rc=$?"
timeout 300 python bench.py --workload c4 --obs-tile 64 --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4_t64.json 2> gpurun_out/bench_c4_t64.err
timeout 300 python bench.py --no-cpu --no-also --obs-to-host-steps 0 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "c3 rc=$?"
MFMARL_OBS_GRID=264 timeout 300 python bench.py --no-cpu --no-also --obs-to-host-steps 0 > gpurun_out/bench_c3_g264.json 2> gpurun_out/bench_c3_g264.err
for f in c4 c4_t64 c3 c3_g264; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$f.json").read().strip().splitlines()[-1])
    print("$f", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], d.get("kernels_alone_ms"), "frac", d["roofline"]["frac"], "e2e %.4g"%d["e2e"]["value"])
except Exception as ex:
    print("$f failed", ex)
PY
done
for pp in fp32 bf16 bf16rows; do
timeout 600 python bench.py --workload play --algo mfq --policy-precision $pp --steps 30 > gpurun_out/bench_play_mfq_$pp.json 2> gpurun_out/bench_play_mfq_$pp.err
timeout 600 python bench.py --workload play --algo mfac --policy-precision $pp --steps 30 > gpurun_out/bench_play_mfac_$pp.json 2> gpurun_out/bench_play_mfac_$pp.err
done
for f in play_mfq_fp32 play_mfq_bf16 play_mfq_bf16rows play_mfac_fp32 play_mfac_bf16 play_mfac_bf16rows; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$f.json").read().strip().splitlines()[-1])
    print("$f", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"])
except Exception as ex:
    print("$f failed", ex, open("gpurun_out/bench_$f.err").read()[-600:])
PY
done
