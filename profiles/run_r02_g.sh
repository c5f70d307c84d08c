#!/bin/bash
# round 2, call G (2 GPUs): data-parallel training test with the NCCL gradient all-reduce, and the bench line at N = 2
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_algo_gpu.py -m gpu -x -q -k data_parallel > gpurun_out/pytest_dp_2gpu.log 2>&1; echo "pytest dp rc=$?" >> gpurun_out/pytest_dp_2gpu.log
tail -5 gpurun_out/pytest_dp_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref n2 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 rc=$?"
tail -3 gpurun_out/bench_n2.err
python - <<PY
import json
for f in ["ref_n2", "n2"]:
    try:
        d=json.loads([l for l in open("gpurun_out/bench_%s.json"%f).read().strip().splitlines() if l.startswith("{")][-1])
        print(f, d.get("n_gpus"), "%.4g"%d["value"], d.get("region_ms"), "e2e %.4g"%d["e2e"]["value"])
        for k,v in d.get("also",{}).items(): print("   also", k, {kk: v[kk] for kk in ("value","ms","bytes","algbw_GBs","mean_matches","n_gpus") if kk in v}, v.get("roofline",{}).get("frac"))
    except Exception as ex:
        print(f, "failed", ex)
PY
