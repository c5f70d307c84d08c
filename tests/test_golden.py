"""Golden fixtures generated from the reference itself (tests/golden/make_golden.py):
CPU: the oracles reproduce them; GPU: the CUDA paths reproduce them.  No engine other than the one under
test runs here, so these hold on the GPU box where /root/reference does not exist."""
import hashlib
import os
import sys

import numpy as np
import pytest

from conftest import REPO

GOLD = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, os.path.join(REPO, "oracle"))


def replay_battle(eng, gold, check_hash=True):
    eng.reset(); eng.add_agents(0, gold["pos0"]); eng.add_agents(1, gold["pos1"])
    deaths = 0
    for s in range(int(gold["steps"])):
        h = hashlib.sha256()
        obs = [eng.get_observation(g) for g in range(2)]
        for v, f in obs:
            h.update(np.ascontiguousarray(v).tobytes()); h.update(np.ascontiguousarray(f).tobytes())
        assert np.array_equal(np.frombuffer(h.digest(), np.uint8), gold["hash_%d" % s]), "obs hash, step %d" % s
        for g in range(2):
            key = "view_%d_%d" % (s, g)
            if key in gold:
                k = len(gold[key])
                assert np.array_equal(obs[g][0][:k].view(np.uint32), gold[key].view(np.uint32))
                assert np.array_equal(obs[g][1][:k].view(np.uint32), gold["feat_%d_%d" % (s, g)].view(np.uint32))
            eng.set_action(g, gold["act_%d_%d" % (s, g)])
        assert eng.step() == bool(gold["done_%d" % s])
        for g in range(2):
            assert np.array_equal(eng.get_reward(g).view(np.uint32), gold["reward_%d_%d" % (s, g)].view(np.uint32))
            al = eng.get_alive(g)
            assert np.array_equal(al, gold["alive_%d_%d" % (s, g)])
            assert np.array_equal(eng.get_pos(g), gold["pos_%d_%d" % (s, g)])
            deaths += int((~al).sum())
        eng.clear_dead()
    return deaths


@pytest.mark.parametrize("name,size", [("battle40_fight", 40), ("battle80_c4", 80)])
def test_oracle_reproduces_reference_battle(name, size):
    from engines import OracleEngine
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    assert replay_battle(OracleEngine(size), gold) >= 5


@pytest.mark.gpu
@pytest.mark.parametrize("name,size", [("battle40_fight", 40), ("battle80_c4", 80)])
def test_cuda_reproduces_reference_battle(name, size):
    from engines import CudaEngine
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    assert replay_battle(CudaEngine(size), gold) >= 5


def test_ising_oracle_reproduces_reference():
    import ising_oracle
    g = np.load(os.path.join(GOLD, "ising20.npz"))
    n, L = int(g["n"]), int(np.sqrt(int(g["n"])))
    spins, Q = g["spins0"].astype(np.int64)[None], np.zeros((1, 5, n, 2))
    for t in range(int(g["steps"])):
        spins, Q, info = ising_oracle.step(spins, Q, float(g["T"]), float(g["lr"]), g["uniforms"][t][None])
        assert np.array_equal(info["action"][0], g["actions"][t])
        assert np.array_equal(info["reward"][0], g["rewards"][t])
        assert info["order"][0] == g["order"][t]
    assert np.array_equal(Q[0].transpose(1, 0, 2), g["Q_final"])


@pytest.mark.gpu
def test_cuda_ising_reproduces_reference():
    import torch
    from mfmarl_b200 import IsingMFQ
    g = np.load(os.path.join(GOLD, "ising20.npz"))
    n = int(g["n"]); L = int(np.sqrt(n))
    m = IsingMFQ(1, L, dtype=torch.float64, lr=float(g["lr"]), spins=torch.from_numpy(g["spins0"].astype(np.int8))[None])
    for t in range(int(g["steps"])):
        m.step(float(g["T"]), uniforms=torch.from_numpy(g["uniforms"][t][None].copy()).cuda())
        assert np.array_equal(m.spins.cpu().numpy().reshape(-1), g["actions"][t])
        assert float(m.reward_sum[0]) == float(g["rewards"][t].sum())
        assert float(m.order_param()[0]) == float(g["order"][t])
    np.testing.assert_allclose(m.q_table()[0].cpu().numpy(), g["Q_final"], rtol=1e-13, atol=1e-15)
