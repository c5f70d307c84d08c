"""The rollout loop of the reference (examples/battle_model/senario_battle.py:41-192 `play`, :196-284 `battle`)
drives the engine through `magent.GridWorld` in a fixed call order and computes the mean action from the
actions with numpy.  This test re-expresses that loop (same calls, same order, same bookkeeping) with small
deterministic stand-in policies and runs it over the CUDA engine and over the C oracle: since the engines
agree bit for bit, every statistic the loop returns must be identical.  GPU."""
import numpy as np
import pytest

from engines import CudaEngine, OracleEngine
from scenarios import StandInPolicy, generate_map_positions

pytestmark = pytest.mark.gpu


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
