set -x
for N in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --no-cpu --obs-to-host-steps 0 > gpurun_out/scale_c3_n$N.json 2> gpurun_out/scale_c3_n$N.err; echo rc=$?; cut -c1-160 gpurun_out/scale_c3_n$N.json; tail -2 gpurun_out/scale_c3_n$N.err
done
python bench.py --no-cpu --obs-to-host-steps 0 > gpurun_out/scale_c3_n1.json 2>/dev/null; cut -c1-160 gpurun_out/scale_c3_n1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --workload c4 --no-cpu --obs-to-host-steps 0 > gpurun_out/scale_c4_n8.json 2> gpurun_out/scale_c4_n8.err; echo rc=$?; cut -c1-160 gpurun_out/scale_c4_n8.json
