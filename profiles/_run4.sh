set -x
timeout 900 python -m pytest tests/test_cuda_ising.py tests/test_golden.py tests/test_algo_gpu.py tests/test_algo_golden.py -m gpu -x -q > gpurun_out/pytest_a.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_a.log; tail -25 gpurun_out/pytest_a.log
for S in 1 25 100; do
timeout 600 python bench.py --workload c5 --steps 100 --sweeps-per-launch $S --no-cpu > gpurun_out/v5_c5_S$S.json 2> gpurun_out/v5_c5_S$S.err; echo "rc=$?"; cut -c1-200 gpurun_out/v5_c5_S$S.json; tail -3 gpurun_out/v5_c5_S$S.err
done
