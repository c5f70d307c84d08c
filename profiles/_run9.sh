set -x
timeout 900 python -m pytest tests/test_cuda_battle_abi.py tests/test_cuda_battle_batched.py tests/test_golden.py tests/test_play_loop.py -m gpu -x -q 2>&1 | tail -3
for W in c3 c3 c4 c4; do
python bench.py --no-cpu --obs-to-host-steps 0 --workload $W > gpurun_out/ab.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1]); print('$W', d['kernels_ms'], d['value'], d['roofline']['frac'], d['e2e']['value'])"
done
python bench.py --no-cpu --obs-to-host-steps 0 --workload c4 --obs-tile 128 > gpurun_out/ab.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1]); print('c4 tile 128', d['kernels_ms'], d['value'])"
