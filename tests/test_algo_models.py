"""PyTorch re-expression of the reference's learners (examples/battle_model/algo/{base,q_learning,ac}.py).  The
reference graphs need TensorFlow 1.x (absent), so these tests pin the arithmetic the TF graph defines -- Q target
(base.py:192-220), masked loss (base.py:94-116), soft update (:84-92), actor-critic losses (ac.py:82-95), discounted
returns (ac.py:139-148) -- against independent numpy / fp32 torch restatements, tolerance 1e-6 relative."""
import os

import numpy as np
import pytest
import torch

from engines import REF_SO, have_ref


class FakeEnv:
    """the three space queries the models make (gridworld.py get_view_space / get_feature_space / get_action_space)"""

    def get_view_space(self, h): return (13, 13, 7)
    def get_feature_space(self, h): return (34,)
    def get_action_space(self, h): return (21,)


def _batch(n, seed=0):
    rng = np.random.RandomState(seed)
    view = rng.rand(n, 13, 13, 7).astype(np.float32)
    feat = rng.rand(n, 34).astype(np.float32)
    prob = rng.dirichlet(np.ones(21), size=n).astype(np.float32)
    return view, feat, prob


def test_value_net_architecture_matches_the_tf_graph():
    from mfmarl_b200.algo import DQN, MFQ
    torch.manual_seed(0)
    il, mf = DQN("il", 0, FakeEnv(), 400, device="cpu"), MFQ("mfq", 0, FakeEnv(), 400, device="cpu")
    count = lambda net: sum(p.numel() for p in net.parameters())
    conv = (3 * 3 * 7 * 32 + 32) + (3 * 3 * 32 * 32 + 32)
    trunk = (9 * 9 * 32 * 256 + 256) + (34 * 32 + 32)
    head = lambda width: (width * 128 + 128) + (128 * 64 + 64) + (64 * 21 + 21)
    assert count(il.eval_net) == conv + trunk + head(288)                                   # base.py:123-183
    assert count(mf.eval_net) == conv + trunk + (21 * 64 + 64) + (64 * 32 + 32) + head(320)
    assert len(il.vars) == 2 * len(list(il.eval_net.parameters()))                         # eval + target scopes


def test_q_target_masked_loss_and_soft_update():
    from mfmarl_b200.algo import MFQ
    torch.manual_seed(1)
    m = MFQ("mfq", 0, FakeEnv(), 400, device="cpu")
    view, feat, prob = _batch(32, 1)
    rng = np.random.RandomState(2)
    rewards, dones = rng.randn(32).astype(np.float32), (rng.rand(32) < 0.3)
    with torch.no_grad():
        tq = m.target_net(torch.from_numpy(view), torch.from_numpy(feat), torch.from_numpy(prob)).numpy()
        eq = m.eval_net(torch.from_numpy(view), torch.from_numpy(feat), torch.from_numpy(prob)).numpy()
    want = rewards + (1.0 - dones) * tq[np.arange(32), eq.argmax(1)] * 0.95                 # base.py:214-219
    got = m.calc_target_q(obs=view, feature=feat, prob=prob, rewards=rewards, dones=dones).numpy()
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-7)
    # act = argmax softmax(Q / T) = argmax Q; eps only renames the temperature (base.py:228-254)
    a1 = m.act(state=[view, feat], prob=prob, eps=1.0)
    a2 = m.act(state=[view, feat], prob=prob, eps=0.05)
    assert a1.dtype == np.int32 and np.array_equal(a1, eq.argmax(1)) and np.array_equal(a1, a2)
    # masked MSE on the taken action (base.py:104-112)
    acts, masks = rng.randint(0, 21, 32).astype(np.int32), (rng.rand(32) < 0.8)
    before = [p.detach().clone() for p in m.eval_net.parameters()]
    tgt_before = [p.detach().clone() for p in m.target_net.parameters()]
    from mfmarl_b200.algo.base import ValueNet
    loss, info = ValueNet.train(m, state=[view, feat], target_q=want.astype(np.float32), prob=prob, acts=acts,
                                masks=masks)
    want_loss = (((want - eq[np.arange(32), acts]) ** 2) * masks).sum() / masks.sum()
    np.testing.assert_allclose(loss, want_loss, rtol=1e-5)
    assert any(not torch.equal(a, b) for a, b in zip(before, m.eval_net.parameters()))
    m.update()                                                                              # t <- tau e + (1 - tau) t
    for t0, t1, e in zip(tgt_before, m.target_net.parameters(), m.eval_net.parameters()):
        np.testing.assert_allclose(t1.detach().numpy(), (0.005 * e.detach() + 0.995 * t0).numpy(), rtol=1e-6, atol=1e-8)


def test_actor_critic_losses_and_returns():
    from mfmarl_b200.algo import MFAC, ActorCritic
    from mfmarl_b200.algo.ac import discounted_returns
    from mfmarl_b200.algo.replay_device import segmented_discounted_returns
    torch.manual_seed(3)
    view, feat, prob = _batch(24, 3)
    rng = np.random.RandomState(4)
    action, ret = rng.randint(0, 21, 24), rng.randn(24).astype(np.float32)
    for cls, p in ((ActorCritic, None), (MFAC, prob)):
        m = cls("m", 0, FakeEnv(), device="cpu", seed=0)
        tv, tf_, ta, tr = (torch.from_numpy(x) for x in (view, feat, action, ret))
        tp = torch.from_numpy(p) if p is not None else None
        with torch.no_grad():
            policy, value = m.net(tv, tf_, tp)
            pg, vf, ent, _ = m.losses(tv, tf_, ta, tr, tp)
        pol, val = policy.numpy().astype(np.float64), value.numpy().astype(np.float64)
        logp = np.log(pol + 1e-6)
        want_pg = -np.mean((ret - val) * logp[np.arange(24), action])                       # ac.py:84-89
        want_vf = 0.1 * np.mean((ret - val) ** 2)
        want_ent = 0.08 * np.mean((pol * logp).sum(1))
        np.testing.assert_allclose([float(pg), float(vf), float(ent)], [want_pg, want_vf, want_ent], rtol=2e-5)
        assert pol.min() >= 1e-10 and abs(pol.sum(1) - 1).max() < 1e-4
        acts = m.act(state=[view, feat])
        assert acts.dtype == np.int32 and acts.shape == (24,) and acts.min() >= 0 and acts.max() < 21
    # discounted returns (ac.py:139-148) and the segmented device form
    r = rng.randn(30).astype(np.float32)
    seg_len = [7, 1, 12, 10]
    seg_last = np.zeros(30, bool); seg_last[np.cumsum(seg_len) - 1] = True
    seg_id = np.repeat(np.arange(4), seg_len)
    boot = rng.randn(4).astype(np.float32)
    want = np.concatenate([discounted_returns(r[seg_id == s], boot[s], 0.95) for s in range(4)])
    keep, ref = None, np.empty(30, np.float32)
    for s in range(4):
        idx = np.where(seg_id == s)[0]
        keep = boot[s]
        for i in idx[::-1]:
            keep = keep * np.float32(0.95) + r[i]; ref[i] = keep
    np.testing.assert_allclose(want, ref, rtol=1e-6)
    got = segmented_discounted_returns(torch.from_numpy(r), torch.from_numpy(seg_last), torch.from_numpy(seg_id),
                                       torch.from_numpy(boot), 0.95).numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-6, atol=1e-6)


def test_batched_actor_critic_update_equals_host_update():
    """the device episode store + segmented returns give the same losses as the reference-shaped host path"""
    from mfmarl_b200.algo import MFAC
    rng = np.random.RandomState(5)
    E, cap, T = 2, 5, 6
    torch.manual_seed(7); host = MFAC("a", 0, FakeEnv(), device="cpu", seed=0)
    torch.manual_seed(7); dev = MFAC("b", 0, FakeEnv(), device="cpu", seed=0, stage_rows=E * cap * T, sub_len=T)
    num = np.array([5, 4], np.int32)
    ids = np.tile(np.arange(cap, dtype=np.int32), (E, 1))
    for t in range(T):
        view = rng.rand(E, cap, 13, 13, 7).astype(np.float32); feat = rng.rand(E, cap, 34).astype(np.float32)
        acts = rng.randint(0, 21, (E, cap)).astype(np.int32); rew = rng.randn(E, cap).astype(np.float32)
        alive = np.ones((E, cap), np.uint8); prob = rng.dirichlet(np.ones(21), size=E).astype(np.float32)
        for e in range(E):
            n = num[e]
            host.flush_buffer(state=[view[e, :n], feat[e, :n]], acts=acts[e, :n], rewards=rew[e, :n], alives=alive[e, :n],
                              ids=[e * 100 + int(i) for i in ids[e, :n]], prob=np.tile(prob[e], (n, 1)))
        tt = torch.from_numpy
        dev.flush_buffer_batched(state=(tt(view), tt(feat)), acts=tt(acts), rewards=tt(rew), alives=tt(alive),
                                 ids=tt(ids), prob=tt(prob), num=tt(num), active=None)
    out_h = host.train(verbose=False)
    out_d = dev.train(verbose=False)
    np.testing.assert_allclose(out_d, out_h, rtol=2e-4, atol=1e-6)
    for a, b in zip(host.net.parameters(), dev.net.parameters()):
        np.testing.assert_allclose(a.detach().numpy(), b.detach().numpy(), rtol=1e-3, atol=2e-5)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("algo", ["mfq", "il", "mfac", "ac"])
def test_runner_trains_one_round_and_saves(algo, tmp_path):
    """train_battle.py's loop body (Runner.run + play + model.train + self-play copy + save/load) for one short round,
    here over the reference CPU engine through the `magent` binding (the GPU twin is in test_algo_gpu.py)."""
    import random
    import magent
    from magent import c_lib
    from mfmarl_b200.algo import spawn_ai, tools
    from mfmarl_b200.senario_battle import play
    random.seed(0); np.random.seed(0); torch.manual_seed(0)
    env = magent.GridWorld("battle", lib=c_lib.load(REF_SO), map_size=40)
    handles = env.get_handles()
    models = [spawn_ai(algo, env, handles[0], algo + "-me", 12, device="cpu"),
              spawn_ai(algo, env, handles[1], algo + "-opponent", 12, device="cpu")]
    runner = tools.Runner(env, handles, 40, 12, models, play, render_every=0, save_every=1, tau=0.01,
                          log_name=algo, log_dir=str(tmp_path / "log"), model_dir=str(tmp_path / "models" / algo),
                          train=True)
    before = [p.detach().clone() for p in models[0].vars]
    info = runner.run(1.0, 0)
    assert set(info["main"]) == {"ave_agent_reward", "total_reward", "kill"}
    assert any(not torch.equal(a, b) for a, b in zip(before, models[0].vars))       # the main model learned something
    models[0].save(str(tmp_path / "ckpt"), 3)
    twin = spawn_ai(algo, env, handles[0], algo + "-twin", 12, device="cpu")
    twin.load(str(tmp_path / "ckpt"), 3)
    for a, b in zip(models[0].vars, twin.vars):
        assert torch.equal(a, b)
    # self-play soft copy (tools.py:566-569): opponent <- (1 - tau) main + tau opponent
    l = [p.detach().clone() for p in models[0].vars]; r = [p.detach().clone() for p in models[1].vars]
    tools.soft_copy(models[1].vars, models[0].vars, 0.01)
    for a, b, c in zip(models[1].vars, l, r):
        np.testing.assert_allclose(a.detach().numpy(), (0.99 * b + 0.01 * c).numpy(), rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("algo", ["mfq", "il", "mfac", "ac"])
def test_bf16_rollout_twins_take_the_engines_bf16_rows(algo):
    """The bf16 twins (base.bf16_rollout_copy, ACNet.bf16_rollout_copy) consume [N, 13, 13, 8] bf16 rows -- the seven
    observation channels plus a zero eighth -- and compute what the fp32 network computes, to bf16 accuracy: the extra
    channel has zero weights, so it is the same function."""
    from mfmarl_b200.algo import spawn_ai
    torch.manual_seed(3)
    m = spawn_ai(algo, FakeEnv(), 0, algo + "-x", 400, device="cpu")
    view, feat, prob = _batch(64, 4)
    view8 = np.zeros((64, 13, 13, 8), np.float32)
    view8[..., :7] = view
    v16, f32, p32 = torch.from_numpy(view8).to(torch.bfloat16), torch.from_numpy(feat), torch.from_numpy(prob)
    if algo in ("mfq", "il"):
        from mfmarl_b200.algo.base import bf16_rollout_copy
        twin = bf16_rollout_copy(m.eval_net)
        with torch.no_grad():
            want = m.eval_net(torch.from_numpy(view), f32, p32 if algo == "mfq" else None)
            got = twin(v16, f32.to(torch.bfloat16), p32.to(torch.bfloat16) if algo == "mfq" else None).float()
        assert got.shape == want.shape and float((got - want).abs().max()) < 0.05 * max(1.0, float(want.abs().max()))
        acts = m.act(state=[v16, f32], prob=p32, eps=1.0)
        assert acts.dtype == torch.int32 and tuple(acts.shape) == (64,)
        assert float((acts == got.argmax(1).to(torch.int32)).float().mean()) == 1.0
    else:
        twin = m.net.bf16_rollout_copy(m.view_space)
        with torch.no_grad():
            want = m.net.policy(torch.from_numpy(view), f32)
            got = twin.policy(v16, f32.to(torch.bfloat16)).float()
        assert float((got - want).abs().max()) < 0.1
        acts = m.act(state=[v16, f32], prob=p32)
        assert acts.dtype == torch.int32 and int(acts.min()) >= 0 and int(acts.max()) < 21
