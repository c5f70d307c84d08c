"""Soak of the persistent Ising kernel's message protocol (K6s): many launches of odd / even sweep counts, lattice
counts that do and do not divide by the slots in flight, three lattice sides; every launch is compared with the
streaming kernel on a sampled subset so that a silent protocol slip (a stale word accepted) would show up as a wrong
lattice, and a hang shows up as the timeout of the caller."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mean-field-multi-agent-reinforcement-learning_b200", "python"))
import torch
from mfmarl_b200 import IsingMFQ

t0 = time.time()
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 45.0
launches = site_steps = 0
rnd = 0
while time.time() - t0 < budget:
    for L, B in ((256, 2048 + rnd % 7), (128, 4100 + rnd % 5), (64, 9000 + rnd % 3), (256, 9), (256, 10), (256, 1)):
        K = (1, 2, 7, 32, 33, 64, 65, 100)[rnd % 8]
        a = IsingMFQ(B, L, seed=rnd)
        nb = min(B, 12)
        b = IsingMFQ(nb, L, seed=rnd, spins=a.spins[:nb].clone())      # same keys: lattice_base 0, lattices 0 .. nb-1
        temps = [0.8 if k % 2 else 1.1 for k in range(K)]
        n_res, _ = a.run(temps, resident=True)
        n_str, _ = b.run(temps, resident=False)
        assert torch.equal(a.spins[:nb], b.spins), (L, B, K)
        assert torch.equal(a.Q[:nb], b.Q), (L, B, K)
        assert torch.equal(n_res[:, :nb], n_str), (L, B, K)
        launches += 1; site_steps += B * L * L * K
        del a, b
    rnd += 1
torch.cuda.synchronize()
print("soak ok: %d resident launches, %.3g site-steps, %d rounds in %.0f s" % (launches, site_steps, rnd, time.time() - t0))
