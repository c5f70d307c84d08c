#!/bin/bash
# round 2, call R (8 GPUs): the default line under torchrun at N = 8 (the driver's scaling step does 1, 2, 4, 8)
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "n8 rc=$?"
tail -3 gpurun_out/bench_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "n4 rc=$?"
python - <<PY
import json
for f in ["n8", "n4"]:
    try:
        d=json.loads([l for l in open("gpurun_out/bench_%s.json"%f).read().strip().splitlines() if l.startswith("{")][-1])
        print(f, d.get("n_gpus"), "%.4g"%d["value"], d.get("region_ms"), "e2e %.4g"%d["e2e"]["value"])
        for k,v in d.get("also",{}).items(): print("   also", k, {kk: v[kk] for kk in ("value","ms","bytes","algbw_GBs","mean_matches") if kk in v}, v.get("roofline",{}).get("frac"), v.get("e2e",{}).get("value"))
    except Exception as ex:
        print(f, "failed", ex)
PY
