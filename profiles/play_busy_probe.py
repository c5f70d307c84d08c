"""How busy is the GPU during play_batched?  Wall time per step against the sum of the kernel times (torch profiler):
if the kernels fill the step, capturing the loop into a CUDA graph has nothing to give."""
import os, sys, json
R = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(R, "mean-field-multi-agent-reinforcement-learning_b200", "python"))
import torch
from torch.profiler import profile, ProfilerActivity
from mfmarl_b200 import BatchedGridWorld
from mfmarl_b200.algo import spawn_ai
from mfmarl_b200.senario_battle import play_batched


class Spaces:
    def get_view_space(self, h): return (13, 13, 7)
    def get_feature_space(self, h): return (34,)
    def get_action_space(self, h): return (21,)


if os.environ.get("CUDNN_BENCHMARK") == "1":
    torch.backends.cudnn.benchmark = True          # let cuDNN time its algorithms instead of using the heuristic
for algo in sys.argv[1:] or ["mfac", "mfq"]:
    dev = torch.device("cuda", 0)
    env = BatchedGridWorld(1024, map_size=40, capacity=64, device=dev, rng="philox", seed=0)
    models = [spawn_ai(algo, Spaces(), g, "%s-%d" % (algo, g), 50, device=dev) for g in range(2)]
    play_batched(env, 0, 5, models, eps=1.0, train=False, left_group=0, obs_dtype=torch.bfloat16)
    torch.cuda.synchronize()
    K = 30
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(); play_batched(env, 1, K, models, eps=1.0, train=False, left_group=0, obs_dtype=torch.bfloat16); t1.record()
    torch.cuda.synchronize()
    wall = t0.elapsed_time(t1) / K
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        play_batched(env, 2, K, models, eps=1.0, train=False, left_group=0, obs_dtype=torch.bfloat16)
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    busy = sum(e.device_time for e in ev) / 1e3 / K
    print("%s bf16rows: %.3f ms per step (events), kernels %.3f ms per step in %d launches per step -> GPU busy %.0f %%"
          % (algo, wall, busy, len(ev) // K, 100 * busy / wall))
    top = {}
    for e in ev:
        top[e.name[:60]] = top.get(e.name[:60], 0) + e.device_time
    for n, t in sorted(top.items(), key=lambda kv: -kv[1])[:8]:
        print("    %6.3f ms/step  %s" % (t / 1e3 / K, n))
    del env, models
