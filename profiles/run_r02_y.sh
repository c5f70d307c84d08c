#!/bin/bash
# round 2, call Y: single-env ABI with the steady step replayed as a CUDA graph
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cuda_battle_abi.py tests/test_render_trace.py -m gpu -x -q > gpurun_out/pytest_abi.log 2>&1; echo "pytest abi rc=$?" >> gpurun_out/pytest_abi.log
tail -5 gpurun_out/pytest_abi.log
timeout 300 python bench.py --workload c2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "c2 rc=$?"
MAGENT_STEP_GRAPH=0 timeout 300 python bench.py --workload c2 > gpurun_out/bench_c2_nograph.json 2> gpurun_out/bench_c2_nograph.err; echo "c2 nograph rc=$?"
timeout 300 python profiles/c2_breakdown.py > gpurun_out/c2_breakdown.txt 2>&1
cat gpurun_out/c2_breakdown.txt | head -30
for f in c2 c2_nograph; do python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$f.json").read().strip().splitlines()[-1])
print("$f", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], d["cpu_baseline"]["value"])
PY
done
