set -x
timeout 600 python -m pytest tests/test_cuda_ising.py -m gpu -x -q > gpurun_out/pytest_ising.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_ising.log; tail -15 gpurun_out/pytest_ising.log
for R in 4 8; do for S in 25 100; do
MFMARL_ISING_RPT=$R timeout 600 python bench.py --workload c5 --steps 100 --sweeps-per-launch $S --no-cpu > gpurun_out/v3_c5_R${R}_S$S.json 2> gpurun_out/v3_c5_R${R}_S$S.err; echo "rc=$?"; cut -c1-200 gpurun_out/v3_c5_R${R}_S$S.json; tail -3 gpurun_out/v3_c5_R${R}_S$S.err
done; done
