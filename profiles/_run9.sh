set -x
for D in 0 1 0 1; do
MFMARL_OBS_DEBUG=$D python bench.py --no-cpu --obs-to-host-steps 0 --steps 100 > gpurun_out/ab.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1]); print('debug=$D c3', d['kernels_ms'], d['value'])"
done
