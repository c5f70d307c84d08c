"""The rollout loop of the reference (examples/battle_model/senario_battle.py:41-192 `play`, :196-284 `battle`)
drives the engine through `magent.GridWorld` in a fixed call order and computes the mean action from the
actions with numpy.  This test re-expresses that loop (same calls, same order, same bookkeeping) with small
deterministic stand-in policies and runs it over the CUDA engine and over the C oracle: since the engines
agree bit for bit, every statistic the loop returns must be identical.  GPU."""
import numpy as np
import pytest

from engines import CudaEngine, OracleEngine
from scenarios import generate_map_positions

pytestmark = pytest.mark.gpu


class StandInPolicy:
    """Duck type of the reference models (algo/base.py:228-254): act(state=[view, feature], prob=, eps=).
    Deterministic given its inputs: attack the first enemy seen in the 8 neighbouring view cells (channel 4),
    otherwise advance towards the other army, the exact move picked by a fixed random linear map of
    (features, mean action) -- so the mean action feeds back into the trajectory as in MF-Q."""

    ADVANCE = {+1: [7, 8, 3, 11], -1: [5, 4, 1, 9]}     # (dx, dy) moves with dx > 0 / dx < 0

    def __init__(self, seed, use_mf, direction):
        rng = np.random.RandomState(seed)
        self.w_feat = rng.randn(34, 4).astype(np.float32)
        self.w_prob = rng.randn(21, 4).astype(np.float32) * (1.0 if use_mf else 0.0)
        self.moves = np.array(self.ADVANCE[direction], np.int32)

    def act(self, state, prob, eps):
        view, feat = state
        assert len(prob) == len(view)
        near = view[:, 5:8, 5:8, 4].reshape(len(view), 9)            # [dy][dx] around the observer
        near = np.delete(near, 4, axis=1)                            # the 8 attack targets, row-major
        q = feat @ self.w_feat + prob.astype(np.float32) @ self.w_prob
        acts = self.moves[np.argmax(q, axis=1)]
        has = near.max(axis=1) > 0
        acts[has] = 13 + np.argmax(near[has] > 0, axis=1)
        return acts.astype(np.int32)


def play_like_reference(eng, models, max_steps):
    """senario_battle.py:96-168, with engine adapters instead of `env` + handles."""
    n_group = 2
    eng.reset()
    left, right = generate_map_positions(eng.map_size)
    eng.add_agents(0, left)
    eng.add_agents(1, right)
    nums = [eng.get_num(g) for g in range(n_group)]
    max_nums = list(nums)
    n_action = [21, 21]
    former_act_prob = [np.zeros((1, 21)), np.zeros((1, 21))]
    mean_rewards, total_rewards = [[], []], [[], []]
    state, ids, acts, rewards, alives = [None] * 2, [None] * 2, [None] * 2, [None] * 2, [None] * 2
    done, step_ct, trace = False, 0, []
    while not done and step_ct < max_steps:
        for i in range(n_group):
            state[i] = list(eng.get_observation(i))
            ids[i] = eng.get_agent_id(i)
        for i in range(n_group):
            former_act_prob[i] = np.tile(former_act_prob[i], (len(state[i][0]), 1))
            acts[i] = models[i].act(state=state[i], prob=former_act_prob[i], eps=1.0)
        for i in range(n_group):
            eng.set_action(i, acts[i])
        done = eng.step()
        for i in range(n_group):
            rewards[i] = eng.get_reward(i)
            alives[i] = eng.get_alive(i)
        for i in range(n_group):   # senario_battle.py:141
            former_act_prob[i] = np.mean(list(map(lambda x: np.eye(n_action[i])[x], acts[i])), axis=0, keepdims=True)
        nums = [eng.get_num(g) for g in range(n_group)]
        for i in range(n_group):
            sum_reward = sum(rewards[i])
            mean_rewards[i].append(sum_reward / nums[i])
            total_rewards[i].append(sum_reward)
        trace.append((ids[0].copy(), acts[0].copy(), rewards[0].copy(), alives[0].copy(), former_act_prob[0].copy()))
        eng.clear_dead()
        step_ct += 1
    for i in range(n_group):
        mean_rewards[i] = sum(mean_rewards[i]) / len(mean_rewards[i])
        total_rewards[i] = sum(total_rewards[i])
    return max_nums, [eng.get_num(g) for g in range(n_group)], mean_rewards, total_rewards, trace


def test_play_loop_statistics_are_identical():
    models = [StandInPolicy(0, True, +1), StandInPolicy(1, False, -1)]   # MF-Q-shaped vs IL-shaped
    out_cuda = play_like_reference(CudaEngine(40), models, max_steps=120)
    out_orac = play_like_reference(OracleEngine(40), models, max_steps=120)
    assert out_cuda[0] == out_orac[0] == [64, 64]
    assert out_cuda[1] == out_orac[1]
    assert out_cuda[2] == out_orac[2] and out_cuda[3] == out_orac[3]     # float sums of identical fp32 rewards
    assert len(out_cuda[4]) == len(out_orac[4])
    kills = 0
    for a, b in zip(out_cuda[4], out_orac[4]):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        kills += int((~a[3]).sum())
    assert sum(out_cuda[1]) < 128 and kills > 0
