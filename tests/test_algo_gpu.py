"""Batched, device-resident rollout (`senario_battle.play_batched`) against the reference-shaped host loop (`play`
through the `magent` binding), both over the CUDA engine, plus one training round of every learner on the GPU."""
import random

import numpy as np
import pytest
import torch

from engines import CUDA_SO

pytestmark = pytest.mark.gpu

ADVANCE = {+1: [7, 8, 3, 11], -1: [5, 4, 1, 9]}


class ExactPolicy:
    """A policy whose decision uses only exactly representable inputs (0/1 view cells, id bits, one mean-action entry
    compared with a threshold no count/n can hit), so the numpy and the torch evaluation agree bit for bit:
    attack the first adjacent enemy, else advance with a move picked from the id bits and the mean action."""

    def __init__(self, direction):
        self.moves_np = np.array(ADVANCE[direction], np.int32)
        self.rows = []

    def act(self, state, prob, eps):
        view, feat = state
        if isinstance(view, torch.Tensor):
            near = view[:, 5:8, 5:8, 4].reshape(view.shape[0], 9)
            near = torch.cat([near[:, :4], near[:, 5:]], dim=1) > 0
            pick = (feat[:, 0] + 2 * feat[:, 1]).to(torch.int64) + (prob[:, 7] > 0.3137).to(torch.int64)
            acts = torch.as_tensor(self.moves_np, device=view.device)[pick % 4]
            first = torch.argmax(near.to(torch.int8), dim=1).to(torch.int32)
            return torch.where(near.any(dim=1), 13 + first, acts).to(torch.int32)
        near = view[:, 5:8, 5:8, 4].reshape(len(view), 9)
        near = np.delete(near, 4, axis=1) > 0
        pick = (feat[:, 0] + 2 * feat[:, 1]).astype(np.int64) + (prob[:, 7] > 0.3137).astype(np.int64)
        acts = self.moves_np[pick % 4]
        return np.where(near.any(axis=1), 13 + np.argmax(near, axis=1), acts).astype(np.int32)

    # recording "learner"
    def flush_buffer(self, **kw):
        self.rows.append({k: np.array(kw[k]) for k in ("ids", "acts", "rewards", "alives")} | {"prob": np.array(kw["prob"][0])})

    def flush_buffer_batched(self, **kw):
        self.rows.append({k: kw[k].cpu().numpy().copy() for k in ("ids", "acts", "rewards", "alives", "prob", "num", "active")})

    def train(self):
        pass


def test_play_batched_equals_host_play_per_environment():
    import magent
    from magent import c_lib
    from mfmarl_b200 import BatchedGridWorld
    from mfmarl_b200.senario_battle import play, play_batched
    steps, E = 70, 3
    # host loop, one environment through the reference binding (group 0 on the left)
    env = magent.GridWorld("battle", lib=c_lib.load(CUDA_SO), map_size=40)
    handles = env.get_handles()
    random.seed(1)
    state = random.getstate(); left = random.randint(0, 1); random.setstate(state)
    host_models = [ExactPolicy(+1 if left == 0 else -1), ExactPolicy(-1 if left == 0 else +1)]
    h_max, h_nums, h_mean, h_total = play(env=env, n_round=0, map_size=40, max_steps=steps, handles=handles,
                                          models=host_models, print_every=1000, eps=1.0, train=True)
    assert sum(h_nums) < 128, "the stand-in policies should produce kills"
    # batched loop, E identical environments on the device
    benv = BatchedGridWorld(E, map_size=40, capacity=64, rng="minstd", seed=0)
    dev_models = [ExactPolicy(+1 if left == 0 else -1), ExactPolicy(-1 if left == 0 else +1)]
    b_max, b_nums, b_mean, b_total = play_batched(benv, 0, steps, dev_models, eps=1.0, train=True, left_group=left)
    for e in range(E):
        assert list(b_max[e]) == list(h_max) and list(b_nums[e]) == list(h_nums)
        np.testing.assert_allclose(b_total[e], np.array(h_total, np.float64), rtol=1e-5, atol=1e-4)
        np.testing.assert_allclose(b_mean[e], np.array(h_mean, np.float64), rtol=1e-5, atol=1e-6)
    assert len(dev_models[0].rows) == len(host_models[0].rows)
    for t, (d, h) in enumerate(zip(dev_models[0].rows, host_models[0].rows)):
        n = len(h["ids"])
        for e in range(E):
            assert d["num"][e] == n and d["active"][e]
            assert np.array_equal(d["ids"][e, :n], h["ids"]), t
            assert np.array_equal(d["acts"][e, :n], h["acts"]), t
            assert np.array_equal(d["rewards"][e, :n].view(np.uint32), h["rewards"].view(np.uint32)), t
            assert np.array_equal(d["alives"][e, :n].astype(bool), h["alives"]), t
            np.testing.assert_allclose(d["prob"][e], h["prob"], rtol=1e-6, atol=1e-7)       # mean action: 1e-6


def test_observe_groups_and_device_state_alias():
    from mfmarl_b200 import BatchedGridWorld
    from scenarios import generate_map_positions, uniform_actions
    E = 5
    env = BatchedGridWorld(E, map_size=40, capacity=64, rng="minstd")
    left, right = generate_map_positions(40)
    env.reset(); env.add_agents(0, left); env.add_agents(1, right)
    rng = np.random.RandomState(0)
    for _ in range(3):
        view, feat = env.observe()
        view, feat = view.clone(), feat.clone()
        (v0, f0), (v1, f1) = env.observe_groups()
        num = env.get_num()
        for e in range(E):
            for g, (vg, fg) in enumerate(((v0, f0), (v1, f1))):
                n = num[e, g]
                assert torch.equal(vg[e, :n], view[e, g, :n]) and torch.equal(fg[e, :n], feat[e, g, :n])
        only1 = env.observe_groups(groups=(1,))
        assert only1[0] is None and torch.equal(only1[1][0], v1)
        assert np.array_equal(env.device_state("num").cpu().numpy(), num)
        assert np.array_equal(env.device_state("id").cpu().numpy(), env.get("id"))
        acts = np.stack([np.stack([uniform_actions(rng, 64) for _ in range(2)]) for _ in range(E)])
        env.step(torch.from_numpy(acts).cuda())


@pytest.mark.parametrize("algo", ["mfq", "il", "mfac", "ac"])
def test_every_learner_trains_a_batched_round_on_the_gpu(algo):
    from mfmarl_b200 import BatchedGridWorld
    from mfmarl_b200.algo import spawn_ai
    from mfmarl_b200.senario_battle import play_batched
    torch.manual_seed(0)

    class Spaces:
        def get_view_space(self, h): return (13, 13, 7)
        def get_feature_space(self, h): return (34,)
        def get_action_space(self, h): return (21,)

    E, steps = 6, 10
    env = BatchedGridWorld(E, map_size=40, capacity=64, rng="philox", seed=3)
    models = [spawn_ai(algo, Spaces(), 0, algo + "-me", steps, device="cuda"),
              spawn_ai(algo, Spaces(), 1, algo + "-opponent", steps, device="cuda")]
    before = [p.detach().clone() for p in models[0].vars]
    max_nums, nums, mean_r, total_r = play_batched(env, 0, steps, models, eps=1.0, train=True, left_group=0)
    assert max_nums.shape == (E, 2) and (max_nums == 64).all() and (nums <= 64).all()
    assert np.isfinite(mean_r).all() and np.isfinite(total_r).all()
    assert any(not torch.equal(a, b) for a, b in zip(before, models[0].vars))
    for p in models[0].vars:
        assert torch.isfinite(p).all()
    # the same round with the policies acting on the engine's bf16 observation rows (bf16 twins, refreshed after the
    # update above); the replay still receives the fp32 rows
    before = [p.detach().clone() for p in models[0].vars]
    max_nums, nums, mean_r, total_r = play_batched(env, 1, steps, models, eps=1.0, train=True, left_group=1,
                                                   obs_dtype=torch.bfloat16)
    assert (max_nums == 64).all() and np.isfinite(mean_r).all() and np.isfinite(total_r).all()
    assert any(not torch.equal(a, b) for a, b in zip(before, models[0].vars))
    assert models[0]._rollout16 is not None and models[0]._rollout16_stale


def test_train_battle_and_battle_scripts_end_to_end(tmp_path):
    """train_battle.py (two rounds, single environment through magent, then two rounds with 8 lock-stepped
    environments) and battle.py on the checkpoints it saved -- the reference's two entry points, same flags."""
    import importlib.util
    import os
    from conftest import PKG

    def load(name):
        spec = importlib.util.spec_from_file_location(name, os.path.join(PKG, "python", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    train, battle = load("train_battle"), load("battle")
    assert train.linear_decay(0, [0, 8, 10], [1, 0.2, 0.1]) == 1 and abs(train.linear_decay(9, [0, 8, 10], [1, 0.2, 0.1]) - 0.15) < 1e-12
    data = str(tmp_path / "data")
    for algo in ("mfq", "mfac"):
        common = ["--algo", algo, "--n_round", "2", "--max_steps", "8", "--data_dir", data]
        runner = train.main(common)
        for tag in ("0", "1"):            # make sure checkpoints exist even if the self-play condition never fired
            runner.models[int(tag)].save(os.path.join(data, "models/%s-%s" % (algo, tag)), 0)
        train.main(common + ["--envs", "8"])
    win = battle.main(["--algo", "mfq", "--oppo", "mfac", "--n_round", "2", "--max_steps", "8", "--idx", "0", "0",
                       "--data_dir", data])
    assert win["main"] + win["opponent"] >= 2


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("algo", ["mfq", "mfac"])
def test_data_parallel_training_keeps_the_ranks_in_step(algo, tmp_path):
    """torchrun, two ranks, each with its own shard of lock-stepped environments (env_base = rank * E) and the optional
    gradient all-reduce: after two rounds both ranks hold bit-identical main-model weights."""
    import os
    import subprocess
    import sys
    from conftest import PKG
    script = os.path.join(PKG, "python", "train_battle.py")
    data = str(tmp_path / "data")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29631", script, "--algo", algo, "--envs", "8", "--n_round", "2",
           "--max_steps", "8", "--save_every", "1", "--data_dir", data]
    env = dict(os.environ, MFMARL_SAVE_FINAL="1")
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    a = torch.load(os.path.join(data, "final_rank0.pt"), map_location="cpu")
    b = torch.load(os.path.join(data, "final_rank1.pt"), map_location="cpu")
    assert len(a) == len(b) > 0
    for x, y in zip(a, b):
        assert torch.equal(x, y)


@pytest.mark.parametrize("use_mf", [True, False])
def test_bf16_rollout_twin_with_the_fused_convolutions_agrees_with_the_plain_path(use_mf):
    """The bf16 rollout twin runs conv + bias + ReLU as ONE cuDNN call (algo.base._conv_bias_relu).  Same weights, same
    bf16 observation rows: its Q values must agree with the unfused bf16 path to bf16 resolution, and its greedy actions
    with the fp32 network on (almost) every row."""
    from mfmarl_b200.algo.base import QNet, bf16_rollout_copy
    torch.manual_seed(5)
    net = QNet((13, 13, 7), (34,), 21, use_mf).cuda()
    with torch.no_grad():
        for conv in (net.conv1, net.conv2):
            conv.bias.uniform_(-0.2, 0.2)                 # (zero at initialisation: make the bias matter)
    N = 4096
    view = torch.zeros((N, 13, 13, 8), device="cuda")
    view[..., :7] = (torch.rand((N, 13, 13, 7), device="cuda") < 0.08).float() * torch.rand((N, 13, 13, 7), device="cuda")
    feat = torch.rand((N, 34), device="cuda")
    prob = torch.softmax(torch.randn((N, 21), device="cuda"), dim=1) if use_mf else None
    twin = bf16_rollout_copy(net)
    assert twin.fused_conv
    v16, f16, p16 = view.to(torch.bfloat16), feat.to(torch.bfloat16), None if prob is None else prob.to(torch.bfloat16)
    with torch.no_grad():
        q_fused = twin(v16, f16, p16).float()
        twin.fused_conv = False
        q_plain = twin(v16, f16, p16).float()
        q32 = net(view[..., :7].contiguous(), feat, prob)
    scale = float(q32.abs().max())
    assert float((q_fused - q_plain).abs().max()) < 0.03 * scale, (float((q_fused - q_plain).abs().max()), scale)
    assert float((q_fused - q32).abs().max()) < 0.05 * scale
    agree = float((q_fused.argmax(1) == q32.argmax(1)).float().mean())
    assert agree > 0.9, agree


def test_play_batched_host_running_ahead_returns_what_the_step_by_step_loop_returns():
    """play_batched looks at the "any environment still playing" flag two steps late so that the host can enqueue ahead
    of the device (senario_battle.py).  With armies that are annihilated after a few steps the round ends early: the
    lagging loop runs at most two more steps, with every environment masked out, and returns exactly what the loop that
    asks after every step returns; the rows handed to the learner are the same, plus all-inactive ones at the end."""
    from mfmarl_b200 import BatchedGridWorld
    from mfmarl_b200.senario_battle import play_batched
    E = 4
    # six attackers around one defender per side pair: group 1 dies within a few steps in every environment
    pos0 = np.array([[10, 10, 0], [11, 10, 0], [12, 10, 0], [10, 12, 0], [11, 12, 0], [12, 12, 0]], np.int32)
    pos1 = np.array([[11, 11, 0]], np.int32)
    out = {}
    for lag in (0, 2):
        env = BatchedGridWorld(E, map_size=40, capacity=64, rng="minstd", seed=0)
        models = [ExactPolicy(+1), ExactPolicy(-1)]
        res = play_batched(env, 0, 50, models, eps=1.0, train=True, left_group=0, positions=(pos0, pos1), host_lag=lag)
        out[lag] = (res, models[0].rows)
    (r0, rows0), (r2, rows2) = out[0], out[2]
    assert len(rows0) < 20, "the round was meant to end early"
    assert len(rows0) <= len(rows2) <= len(rows0) + 2
    for a, b in zip(r0, r2):
        assert np.array_equal(a, b)
    for d0, d2 in zip(rows0, rows2):
        for k in ("ids", "acts", "rewards", "alives", "prob", "num", "active"):
            assert np.array_equal(d0[k], d2[k]), k
    for extra in rows2[len(rows0):]:
        assert not extra["active"].any()
