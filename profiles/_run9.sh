set -x
for D in 0 2 3 0 2; do
MFMARL_OBS_DEBUG=$D python bench.py --no-cpu --obs-to-host-steps 0 --steps 100 > gpurun_out/dbg_$D.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/dbg_$D.json').read().strip().splitlines()[-1]); print('debug=$D', d['kernels_ms'])"
done
python profiles/write_probe.py
MFMARL_OBS_DEBUG=2 timeout 300 python -m pytest tests/test_cuda_battle_batched.py -m gpu -x -q 2>&1 | tail -2
