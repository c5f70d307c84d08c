set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log; tail -25 gpurun_out/pytest_gpu.log
python bench.py --no-cpu > gpurun_out/bench_c3_b.json 2> gpurun_out/bench_c3_b.err; echo rc=$?; cut -c1-300 gpurun_out/bench_c3_b.json
python -c "
import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo rc=$?; tail -3 gpurun_out/smoke.log
