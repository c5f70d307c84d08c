"""Does a streaming launch train on a second stream (lattices the resident kernel does not take) pay off?  The resident
kernel keeps 112 of 148 SMs at ~50 % issue utilisation and uses no HBM bandwidth; the streaming kernel is HBM-bound."""
import os, sys, ctypes
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mean-field-multi-agent-reinforcement-learning_b200", "python"))
import torch
from mfmarl_b200 import IsingMFQ

B, L, S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192, 256, 100
for frac in (0.0, 0.1, 0.15, 0.2, 0.25, 0.3):
    Bs = int(B * frac) // 8 * 8
    Br = B - Bs
    res = IsingMFQ(Br, L, seed=13)
    stream_part = IsingMFQ(Bs, L, seed=13, lattice_base=Br) if Bs else None
    side = torch.cuda.Stream()
    temps = [0.8] * S

    def go():
        main = torch.cuda.current_stream()
        side.wait_stream(main)
        res.run(temps, resident=True)                       # resident first: its clusters need whole SMs
        if stream_part is not None:
            with torch.cuda.stream(side):
                for k in range(S):
                    stream_part.step(0.8, stats=False)
            main.wait_stream(side)

    go(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); go(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("stream fraction %.2f (%5d resident + %5d streaming lattices): %.1f ms for %d sweeps -> %.3e site-steps/s"
          % (frac, Br, Bs, ms, S, B * L * L * S / ms * 1e3), flush=True)
    del res, stream_part
