set -x
timeout 600 python -m pytest tests/test_render_trace.py tests/test_algo_gpu.py -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/bench_c3_c.json 2> gpurun_out/bench_c3_c.err; echo rc=$?; tail -2 gpurun_out/bench_c3_c.err
for A in mfq mfac; do
python bench.py --workload play --algo $A --steps 50 --envs 1024 > gpurun_out/bench_play_$A.json 2> gpurun_out/bench_play_$A.err; echo rc=$?; cut -c1-250 gpurun_out/bench_play_$A.json; tail -3 gpurun_out/bench_play_$A.err
done
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err; echo rc=$?; cut -c1-300 gpurun_out/bench_ref.json
