set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --workload c4 --no-cpu --obs-to-host-steps 0 > gpurun_out/bench_c4_v5.json 2>gpurun_out/bench_c4_v5.err; echo rc=$?
python bench.py --workload c4 --no-cpu --obs-to-host-steps 0 --pipeline 1 > gpurun_out/bench_c4_v5_p1.json 2>gpurun_out/bench_c4_v5_p1.err; echo rc=$?
python bench.py --no-cpu --obs-to-host-steps 0 --pipeline 1 > gpurun_out/bench_c3_v5_p1.json 2> gpurun_out/bench_c3_v5_p1.err; echo rc=$?
python - <<'PY'
import json
for f in ['bench_c3_v5_p1','bench_c4_v5','bench_c4_v5_p1']:
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, '%.4e'%d['value'], d['kernels_ms']['k_obs'], d['kernels_ms']['k_step'], d.get('kernels_alone_ms'), d['roofline']['frac'], '%.4e'%d['e2e']['value'])
PY
