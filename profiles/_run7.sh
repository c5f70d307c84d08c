set -x
nvidia-smi -L | head -8
for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu --obs-to-host-steps 0 > gpurun_out/scale_c3_n$N.json 2> gpurun_out/scale_c3_n$N.err; echo rc=$?; cut -c1-200 gpurun_out/scale_c3_n$N.json; tail -2 gpurun_out/scale_c3_n$N.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --workload c5 --steps 100 --sweeps-per-launch 100 --no-cpu > gpurun_out/scale_c5_n8.json 2> gpurun_out/scale_c5_n8.err; echo rc=$?; cut -c1-200 gpurun_out/scale_c5_n8.json; tail -2 gpurun_out/scale_c5_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --workload c4 --no-cpu --obs-to-host-steps 0 > gpurun_out/scale_c4_n8.json 2> gpurun_out/scale_c4_n8.err; echo rc=$?; cut -c1-200 gpurun_out/scale_c4_n8.json; tail -2 gpurun_out/scale_c4_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --impl reference --steps 3 --warmup 3 > gpurun_out/scale_ref_n8.json 2> gpurun_out/scale_ref_n8.err; echo rc=$?; cut -c1-200 gpurun_out/scale_ref_n8.json
