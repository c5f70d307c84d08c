set -x
timeout 900 python -m pytest tests/test_algo_gpu.py tests/test_cuda_battle_abi.py -m gpu -x -q 2>&1 | tail -3
python bench.py --workload c2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo rc=$?; cut -c1-250 gpurun_out/bench_c2.json; tail -2 gpurun_out/bench_c2.err
for PP in fp32 tf32 bf16; do
python bench.py --workload play --algo mfq --steps 50 --envs 1024 --policy-precision $PP > gpurun_out/bench_play_mfq_$PP.json 2> gpurun_out/bench_play_mfq_$PP.err; echo rc=$?; cut -c1-160 gpurun_out/bench_play_mfq_$PP.json; tail -2 gpurun_out/bench_play_mfq_$PP.err
done
python bench.py --workload c5 > gpurun_out/bench_c5_default.json 2> gpurun_out/bench_c5_default.err; echo rc=$?; cut -c1-200 gpurun_out/bench_c5_default.json
