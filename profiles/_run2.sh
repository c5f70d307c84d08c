set -x
timeout 600 python -m pytest tests/test_cuda_ising.py tests/test_play_loop.py tests/test_golden.py -m gpu -x -q > gpurun_out/pytest_ising.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_ising.log; tail -15 gpurun_out/pytest_ising.log
for S in 1 25 100; do
timeout 600 python bench.py --workload c5 --steps 100 --sweeps-per-launch $S --no-cpu > gpurun_out/v2_c5_S$S.json 2> gpurun_out/v2_c5_S$S.err; echo "rc=$?"; cut -c1-400 gpurun_out/v2_c5_S$S.json; tail -3 gpurun_out/v2_c5_S$S.err
done
