set -x
python bench.py --workload c5 --envs 2048 --steps 50 --warmup 3 --sweeps-per-launch 25 --no-cpu > gpurun_out/v4_c5_small.json 2>gpurun_out/v4_c5_small.err; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:k_ising_resident -s 2 -c 1 -o gpurun_out/ising_resident_v4 -f python bench.py --workload c5 --envs 2048 --steps 50 --warmup 3 --sweeps-per-launch 25 --no-cpu > gpurun_out/ncu_c5_v4.log 2>&1
python -m pytest tests/test_algo_golden.py tests/test_play_loop.py -m gpu -x -q 2>&1 | tail -5
ls -la gpurun_out | tail -3
