"""Ising kernel probe: streaming (one launch per sweep) vs resident (K sweeps per launch) throughput."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mean-field-multi-agent-reinforcement-learning_b200", "python"))
import torch
from mfmarl_b200 import IsingMFQ

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
L = int(sys.argv[3]) if len(sys.argv) > 3 else 256
mode = sys.argv[4] if len(sys.argv) > 4 else "both"
for resident in ([False, True] if mode == "both" else [mode == "resident"]):
    m = IsingMFQ(B, L, seed=13)
    m.run([0.8] * 3, resident=resident)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    m.run([0.8] * K, resident=resident)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print("%s B=%d L=%d K=%d ms/sweep %.3f site-steps/s %.3e algorithmic GB/s %.0f order %.3f" % (
        "resident " if resident else "streaming", B, L, K, ms, B * L * L / ms * 1e3, B * L * L * 14 / ms / 1e6,
        float(m.order_param().mean())))
    del m
