#!/bin/bash
# round 2, call Z (8 GPUs): the default line under torchrun at N = 8, 4, 2 on the final tree (K6s, step graph)
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for N in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29530 + N)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "n$N rc=$?"
tail -2 gpurun_out/bench_n$N.err
done
python - <<PY
import json
for f in ["n8", "n4", "n2"]:
    try:
        d=json.loads([l for l in open("gpurun_out/bench_%s.json"%f).read().strip().splitlines() if l.startswith("{")][-1])
        print(f, d.get("n_gpus"), "%.4g"%d["value"], d.get("region_ms"), "e2e %.4g"%d["e2e"]["value"])
        for k,v in d.get("also",{}).items(): print("   also", k, {kk: v[kk] for kk in ("value","ms","bytes","algbw_GBs","mean_matches") if kk in v}, v.get("roofline",{}).get("frac"), v.get("e2e",{}).get("value"))
    except Exception as ex:
        print(f, "failed", ex)
PY
