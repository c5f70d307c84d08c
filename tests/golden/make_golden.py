"""Generates the golden fixtures of tests/golden/ from the REFERENCE itself (run in the dev container, where
/root/reference exists):

  battle40_fight.npz   the reference C++ engine (oracle/_ref/libmagent_ref.so, OMP_NUM_THREADS=1) stepped for
                       120 steps on the seeded fight stream: per step the actions, rewards, alive flags,
                       positions after the step, a SHA-256 of both groups' views + features, and the full
                       observation arrays at steps 0, 40 and 80.
  battle80_c4.npz      the same for the 80x80 512 v 512 placement (30 steps, hashes + step 0 group-0 slice).
  ising20.npz          the unmodified reference Ising classes (under oracle/ising_ref_shim) for 40 steps of the
                       main_MFQ_Ising.py loop body at tau = 0.8 with injected uniforms: initial spins, uniforms,
                       and per step actions, rewards, order parameter; final Q.

    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import conftest  # noqa: F401,E402  (import paths, OMP_NUM_THREADS=1)
from engines import RefEngine  # noqa: E402
from scenarios import c4_positions, fight_actions, generate_map_positions  # noqa: E402


def obs_hash(eng):
    h = hashlib.sha256()
    for g in range(2):
        v, f = eng.get_observation(g)
        h.update(np.ascontiguousarray(v).tobytes()); h.update(np.ascontiguousarray(f).tobytes())
    return np.frombuffer(h.digest(), dtype=np.uint8)


def battle(map_size, pos, steps, seed, keep_obs_at, keep_slice=None):
    eng = RefEngine(map_size)
    eng.reset(); eng.add_agents(0, pos[0]); eng.add_agents(1, pos[1])
    rng = np.random.RandomState(seed)
    out = {"map_size": map_size, "seed": seed, "pos0": pos[0], "pos1": pos[1]}
    for s in range(steps):
        out["hash_%d" % s] = obs_hash(eng)
        if s in keep_obs_at:
            for g in range(2):
                v, f = eng.get_observation(g)
                if keep_slice:
                    v, f = v[:keep_slice], f[:keep_slice]
                out["view_%d_%d" % (s, g)], out["feat_%d_%d" % (s, g)] = v, f
        for g in range(2):
            a = fight_actions(rng, eng.get_pos(g), map_size)
            out["act_%d_%d" % (s, g)] = a
            eng.set_action(g, a)
        out["done_%d" % s] = np.array(eng.step())
        for g in range(2):
            out["reward_%d_%d" % (s, g)] = eng.get_reward(g)
            out["alive_%d_%d" % (s, g)] = eng.get_alive(g)
            out["pos_%d_%d" % (s, g)] = eng.get_pos(g)
        eng.clear_dead()
    out["steps"] = steps
    return out


def ising(n=400, T=0.8, steps=40, lr=0.1):
    sys.path.insert(0, os.path.join(conftest.REPO, "oracle", "ising_ref_shim"))
    sys.path.insert(0, "/root/reference")
    from examples.ising_model.multiagent.environment import IsingMultiAgentEnv
    import examples.ising_model as im
    sc = im.load("Ising.py").Scenario()
    np.random.seed(13)
    env = IsingMultiAgentEnv(world=sc.make_world(num_agents=n, agent_view=1), reset_callback=sc.reset_world,
                             reward_callback=sc.reward, observation_callback=sc.observation,
                             done_callback=sc.done)
    obs = np.stack(env.reset())
    out = {"n": n, "T": T, "lr": lr, "steps": steps, "spins0": env.world.global_state.astype(np.int8).copy()}
    Q = np.zeros((n, 5, 2))
    rng = np.random.RandomState(2024)
    U, A, R, O = [], [], [], []
    for t in range(steps):
        u = rng.random_sample(n)
        action = np.zeros(n, dtype=np.int32)
        for i in range(n):
            s = np.count_nonzero(obs[i] == 1)
            vals = [np.exp(Q[i, s, k] / T) for k in range(2)]
            denom = 0
            for v in vals:
                denom += v
            cdf = np.array([v / denom for v in vals]).cumsum()
            cdf /= cdf[-1]
            action[i] = cdf.searchsorted(u[i], side="right")
        obs_, reward, done, order, ups, downs = env.step(np.expand_dims(action, 1))
        for i in range(n):
            s = np.count_nonzero(obs[i] == 1)
            Q[i, s, action[i]] = Q[i, s, action[i]] + lr * (float(reward[i][0]) - Q[i, s, action[i]])
        obs = np.stack(obs_)
        U.append(u); A.append(action.astype(np.int8)); R.append(np.array(reward).reshape(-1)); O.append(order)
    out.update(uniforms=np.array(U), actions=np.array(A), rewards=np.array(R), order=np.array(O), Q_final=Q)
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "battle40_fight.npz"),
                        **battle(40, generate_map_positions(40), 120, 7, keep_obs_at=(0, 40, 80)))
    np.savez_compressed(os.path.join(HERE, "battle80_c4.npz"),
                        **battle(80, c4_positions(), 30, 8, keep_obs_at=(0,), keep_slice=64))
    np.savez_compressed(os.path.join(HERE, "ising20.npz"), **ising())
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
