"""Ising kernel probe: throughput in the disordered (fresh) and ordered (converged) phase."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mean-field-multi-agent-reinforcement-learning_b200", "python"))
import torch
from mfmarl_b200 import IsingMFQ

B, L = int(sys.argv[1]) if len(sys.argv) > 1 else 2048, 256
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
pre = int(sys.argv[3]) if len(sys.argv) > 3 else 0
m = IsingMFQ(B, L, seed=13)
for _ in range(3 + pre):
    m.step(0.8)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(K):
    m.step(0.8)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print("B=%d L=%d pre=%d ms/sweep %.3f site-steps/s %.3e algorithmic GB/s %.0f order %.3f" % (
    B, L, pre, ms, B * L * L / ms * 1e3, B * L * L * 14 / ms / 1e6, float(m.order_param().mean())))
