"""k_obs alone, fp32 rows against bf16 NHWC-8 rows (mfb_observe_groups / mfb_observe_groups_bf16), at the C3 and C4 shapes:
agents observed per second and bytes written per second.     python profiles/obs_rows_probe.py"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "mean-field-multi-agent-reinforcement-learning_b200", "python"))
from mfmarl_b200 import BatchedGridWorld  # noqa: E402
from mfmarl_b200.scenarios import c4_positions, generate_map_positions  # noqa: E402

for name, E, ms, cap, pos in (("C3 40x40 64v64 x4096", 4096, 40, 64, generate_map_positions(40)),
                              ("C4 80x80 512v512 x128", 128, 80, 512, c4_positions())):
    env = BatchedGridWorld(E, map_size=ms, capacity=cap, rng="philox")
    env.reset(); env.add_agents(0, pos[0]); env.add_agents(1, pos[1])
    n = int(env.get_num().sum())
    for dtype, row_bytes in ((torch.float32, 4732 + 136), (torch.bfloat16, 2704 + 136)):
        for _ in range(3):
            env.observe_groups(dtype=dtype)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 50
        t0.record()
        for _ in range(reps):
            env.observe_groups(dtype=dtype)
        t1.record()
        torch.cuda.synchronize()
        ms_launch = t0.elapsed_time(t1) / reps
        print("%-24s %-9s %8.4f ms per launch  %7.3e agents/s  %6.0f GB/s written" % (
            name, str(dtype).split(".")[1], ms_launch, n / (ms_launch * 1e-3), n * row_bytes / (ms_launch * 1e-3) / 1e9), flush=True)
    del env
    torch.cuda.empty_cache()
