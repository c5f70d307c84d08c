set -x
cd mean-field-multi-agent-reinforcement-learning_b200/python
for A in mfq mfac; do
/usr/bin/time -v python train_battle.py --algo $A --envs 256 --n_round 3 --max_steps 100 --data_dir /tmp/tb_$A > /root/repo/gpurun_out/train_$A.log 2>&1; echo rc=$?
grep -n "ROUND\|INFO\] {\|LOSS\|Elapsed\|Maximum resident" /root/repo/gpurun_out/train_$A.log | tail -14
done
timeout 600 python -m pytest /root/repo/tests/test_algo_gpu.py -m gpu -x -q 2>&1 | tail -3
