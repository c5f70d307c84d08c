set -x
timeout 900 python -m pytest tests/test_cuda_battle_abi.py tests/test_cuda_battle_batched.py tests/test_golden.py tests/test_play_loop.py -m gpu -x -q 2>&1 | tail -5
python bench.py --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c3_d.json 2> gpurun_out/bench_c3_d.err; echo rc=$?
python bench.py --no-cpu --obs-to-host-steps 0 --workload c4 > gpurun_out/bench_c4_d.json 2> gpurun_out/bench_c4_d.err; echo rc=$?
python - <<'PY'
import json
for f in ['bench_c3_d','bench_c4_d']:
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, '%.4e'%d['value'], d['kernels_ms'], d['roofline']['frac'], '%.4e'%d['e2e']['value'])
PY
