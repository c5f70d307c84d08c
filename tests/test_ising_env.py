"""The Ising ENVIRONMENT interface over the CUDA kernels: `mfi_env_step` against numpy, and the
reference script's loop over python/examples/ising_model against the text the UNMODIFIED main_MFQ_Ising.py printed
over the UNMODIFIED reference environment (tests/golden/ising_main_*.txt, made by make_golden_ising_main.py)."""
import ctypes
import io
import os
import runpy
import sys
from contextlib import redirect_stdout

import numpy as np
import pytest

from conftest import REPO

GOLDEN = os.path.join(REPO, "tests", "golden")
REF_SCRIPT = "/root/reference/main_MFQ_Ising.py"


def script_loop(argv):
    """What main_MFQ_Ising.py:11-172 does, restated with the same numpy calls in the same order (the script itself is
    not on the GPU box): seed 13, Scenario.make_world + env.reset draw the spins, one np.random.choice(2, 1, p) per
    agent per step, one np.random.choice(n, k, replace=False) per step.  Returns the printed lines."""
    import argparse
    from examples.ising_model.multiagent.environment import IsingMultiAgentEnv
    import examples.ising_model as ising_model
    ap = argparse.ArgumentParser()
    ap.add_argument('-n', type=int, default=100); ap.add_argument('-t', type=float, default=1)
    ap.add_argument('-ts', type=int, default=10000); ap.add_argument('-lr', type=float, default=0.1)
    ap.add_argument('-dr', type=float, default=0.99); ap.add_argument('-dg', type=int, default=2000)
    ap.add_argument('-ac', type=float, default=1.0)
    a = ap.parse_args(argv)
    np.random.seed(13)
    sc = ising_model.load('Ising.py').Scenario()
    env = IsingMultiAgentEnv(world=sc.make_world(num_agents=a.n, agent_view=1), reset_callback=sc.reset_world,
                             reward_callback=sc.reward, observation_callback=sc.observation, done_callback=sc.done)
    n, n_actions = env.n, env.action_space[0].n
    assert env.observation_space[0].n == 4 and n_actions == 2
    target = np.array([[2, -2], [1, -1], [0, 0], [-1, 1], [-2, 2]])
    lines = []
    obs = np.stack(env.reset())
    Q = np.zeros((n, 5, n_actions))
    max_order, max_step, o_up, o_down, stagnant, cur_t = 0.0, 0, 0, 0, 0, 0.3
    for t in range(a.ts):
        if t % a.dg == 0:
            cur_t *= a.dr
        if cur_t < a.t:
            cur_t = a.t
        action = np.zeros(n, dtype=np.int32)
        for i in range(n):
            s = np.count_nonzero(obs[i] == 1)
            vals, denom = [], 0
            for k in range(n_actions):
                v = np.exp(Q[i, s, k] / cur_t)
                vals.append(v); denom += v
            action[i] = np.random.choice(n_actions, 1, p=[x / denom for x in vals])[0]
        obs_, reward, done, order, ups, downs = env.step(np.expand_dims(action, axis=1))
        obs_ = np.stack(obs_)
        mse = 0
        for i in np.random.choice(n, int(a.ac * n), replace=False):
            s = np.count_nonzero(obs[i] == 1)
            Q[i, s, action[i]] = Q[i, s, action[i]] + a.lr * (reward[i] - Q[i, s, action[i]])
            mse += np.power((Q[i, s, action[i]] - target[s, action[i]]), 2)
        mse /= n
        obs = obs_
        if order > max_order:
            max_order, max_step, o_up, o_down = order, t, ups, downs
            lines.append("+++++++++++++++++++++++++++++")
        stagnant = stagnant + 1 if abs(max_order - order) < 0.001 else 0
        if stagnant == 500 or t > a.ts:
            break
        lines.append('E: %d/%d, reward = %f, mse = %f, Order = %f, Up = %d, Down = %d'
                     % (0, t, float(np.sum(reward)), float(np.asarray(mse).reshape(-1)[0]), order, ups, downs))
    lines.append('Episode: %d, MaxO = %f at %d (%d/%d)' % (0, max_order, max_step, o_up, o_down))
    return lines


CASES = [("ising_main_n400_t0.8_ts50.txt", ["-n", "400", "-t", "0.8", "-ts", "50"]),
         ("ising_main_n100_t0.25_ts40_ac0.6.txt", ["-n", "100", "-t", "0.25", "-ts", "40", "-ac", "0.6", "-dg", "7"])]


@pytest.mark.gpu
@pytest.mark.filterwarnings("ignore::DeprecationWarning")
@pytest.mark.parametrize("golden,argv", CASES)
def test_script_loop_over_the_cuda_env_prints_what_the_reference_printed(golden, argv):
    want = open(os.path.join(GOLDEN, golden)).read().splitlines()
    got = script_loop(argv)
    assert got == want


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(REF_SCRIPT), reason="the reference script is only present in the dev container")
@pytest.mark.parametrize("golden,argv", CASES)
def test_unmodified_reference_script_runs_over_the_cuda_env(golden, argv, tmp_path, monkeypatch):
    """main_MFQ_Ising.py itself, byte for byte, with python/ ahead of the reference tree on sys.path."""
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(sys, "argv", [REF_SCRIPT] + argv)
    for name in [m for m in sys.modules if m == "examples" or m.startswith("examples.")]:
        del sys.modules[name]
    buf = io.StringIO()
    with redirect_stdout(buf):
        runpy.run_path(REF_SCRIPT, run_name="__main__")
    assert buf.getvalue().splitlines() == open(os.path.join(GOLDEN, golden)).read().splitlines()


@pytest.mark.gpu
@pytest.mark.parametrize("B,L", [(1, 3), (2, 20), (3, 33), (2, 256)])
def test_mfi_env_step_matches_numpy(B, L):
    """spin <- [action > 0] (environment.py:112-114), reward 0.5 sigma_i sum sigma_j on the NEW lattice (Ising.py:101-111),
    observation = neighbour spins by ascending flat index (Ising.py:113-118), n_up (core.py:106-110)."""
    import torch
    from mfmarl_b200.lib import check, load_library
    lib = load_library()
    lib.mfi_env_step.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 6
    rng = np.random.RandomState(L)
    N = L * L
    spins = torch.from_numpy(rng.randint(0, 2, size=(B, N)).astype(np.int8)).cuda()
    obs = torch.zeros((B, N, 4), dtype=torch.uint8, device="cuda")
    rew = torch.zeros((B, N), dtype=torch.float32, device="cuda")
    nup = torch.zeros((B,), dtype=torch.int32, device="cuda")
    idx = np.arange(N).reshape(L, L)
    nb = np.sort(np.stack([np.roll(idx, 1, 0), np.roll(idx, -1, 0), np.roll(idx, 1, 1), np.roll(idx, -1, 1)], -1)
                 .reshape(N, 4), axis=1)
    for t in range(4):
        acts = rng.randint(-2, 3, size=(B, N)).astype(np.int32) if t else None
        d_acts = torch.from_numpy(acts).cuda() if t else None
        before = spins.cpu().numpy()
        check(lib.mfi_env_step(B, L, spins.data_ptr(), d_acts.data_ptr() if t else None, obs.data_ptr(),
                               rew.data_ptr(), nup.data_ptr(), None))
        torch.cuda.synchronize()
        want = (acts > 0).astype(np.int8) if t else before
        assert np.array_equal(spins.cpu().numpy(), want)
        sig = 2.0 * want - 1.0
        assert np.array_equal(obs.cpu().numpy(), want[:, nb].astype(np.uint8))
        assert np.array_equal(rew.cpu().numpy(), (0.5 * sig * sig[:, nb].sum(-1)).astype(np.float32))
        assert np.array_equal(nup.cpu().numpy(), want.sum(1))


@pytest.mark.gpu
def test_env_object_surface():
    """attributes and per-agent callbacks the reference classes expose (environment.py:13-47, core.py:57-97)"""
    import examples.ising_model as im
    from examples.ising_model.multiagent.environment import IsingMultiAgentEnv
    np.random.seed(3)
    sc = im.load("Ising.py").Scenario()
    world = sc.make_world(num_agents=49, agent_view=1)
    env = IsingMultiAgentEnv(world=world, reset_callback=sc.reset_world, reward_callback=sc.reward,
                             observation_callback=sc.observation, done_callback=sc.done)
    assert env.n == 49 and world.shape_size == 7 and len(world.policy_agents) == 49 and not world.scripted_agents
    obs = env.reset()
    assert len(obs) == 49 and obs[0].shape == (4,) and obs[0].dtype == np.float64
    acts = np.random.randint(0, 2, size=(49, 1)).astype(np.int32)
    obs_n, reward_n, done_n, order, ups, downs = env.step(acts)
    assert ups + downs == 49 and order == abs(ups - downs) / 49.0 and done_n == [order == 1.0] * 49
    assert np.array_equal(world.global_state.flatten(), acts[:, 0].astype(np.float64))
    for i in (0, 6, 24, 48):                  # the per-agent callbacks agree with the batched answers
        agent = world.agents[i]
        assert agent.state.id == i and agent.state.spin == acts[i, 0]
        assert np.array_equal(sc.observation(agent, world), obs_n[i])
        assert reward_n[i].shape == (1,) and sc.reward(agent, world)[0] == reward_n[i][0]
    ones = np.ones((49, 1), np.int32)
    assert env.step(ones)[2] == [True] * 49 and world.order_param == 1.0
    with pytest.raises(NotImplementedError):
        IsingMultiAgentEnv(world=world, reset_callback=sc.reset_world, reward_callback=lambda a, w: 0.0)
