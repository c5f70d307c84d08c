set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --workload c2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo rc=$?; python -c "
import json; d=json.loads(open('gpurun_out/bench_c2.json').read().strip().splitlines()[-1]); print('c2', d['value'], d['ms_per_step'], d.get('cpu_baseline'))"
python -c "
import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
