"""Rows either side of the hot path against fixtures produced by the REFERENCE's own Python
(tests/golden/make_golden_algo.py): the rollout loop `play` (senario_battle.py:41-192) and the replay buffers
(algo/tools.py:26-362)."""
import importlib.util
import os

import numpy as np
import pytest

from engines import CUDA_SO, REF_SO, have_ref

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def _gen():
    spec = importlib.util.spec_from_file_location("make_golden_algo", os.path.join(GOLD, "make_golden_algo.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _check_play(lib_path):
    from mfmarl_b200 import senario_battle
    gen = _gen()
    out, models = gen.golden_play(senario_battle.play, lib_path)
    got = gen.pack_play(out, models)
    want = np.load(os.path.join(GOLD, "play_round.npz"))
    assert set(got) == set(want.files)
    assert int(want["steps"]) == 60 and int(want["trained"]) == 1
    assert want["nums"].sum() < 128                      # the round has kills
    for k in want.files:
        assert np.array_equal(np.asarray(got[k]), want[k]), k


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_play_loop_over_reference_engine_matches_reference_play():
    """my play() == the reference's play(): same engine (unmodified reference C++), same stand-in models."""
    _check_play(REF_SO)


@pytest.mark.gpu
def test_play_loop_over_cuda_engine_matches_reference_play():
    """my play() over the CUDA engine == the reference's play() over the reference engine, every flushed row."""
    _check_play(CUDA_SO)


def test_host_replay_buffers_match_reference_tools():
    from mfmarl_b200.algo import tools
    gen = _gen()
    want = np.load(os.path.join(GOLD, "tools_replay.npz"))
    obs_shape, feat_shape, act_n = (3, 3, 2), (5,), 4
    for use_mean in (False, True):
        np.random.seed(11)
        mg = tools.MemoryGroup(obs_shape, feat_shape, act_n, max_len=150, batch_size=16, sub_len=12, use_mean=use_mean)
        for rnd in range(3):
            for kw in gen.synthetic_stream(100 + rnd, 9, 10, obs_shape, feat_shape, act_n):
                mg.push(**kw)
            mg.tight()
        tag = "mf" if use_mean else "il"
        assert mg.nb_entries == int(want[tag + "_len"]) == 150
        for name, arr in (("obs0", mg.obs0.pull()), ("act", mg.actions.pull()), ("rew", mg.rewards.pull()),
                          ("term", mg.terminals.pull()), ("mask", mg.masks.pull())):
            assert np.array_equal(arr, want["%s_%s" % (tag, name)]), name
        for k in range(2):
            for i, arr in enumerate(mg.sample()):
                assert np.array_equal(np.asarray(arr), want["%s_sample%d_%d" % (tag, k, i)]), (k, i)
    np.random.seed(12)
    eb = tools.EpisodesBuffer(use_mean=True)
    for kw in gen.synthetic_stream(200, 7, 8, obs_shape, feat_shape, act_n):
        eb.push(**kw)
    eps = list(eb.episodes())
    assert np.array_equal(np.array(list(eb.buffer.keys())), want["eb_order"])
    assert np.array_equal(np.array([len(e.rewards) for e in eps]), want["eb_lens"])
    assert np.array_equal(np.array([e.terminal for e in eps]), want["eb_terminal"])
    assert np.array_equal(np.array(eps[0].rewards), want["eb_rewards0"])
    assert np.array_equal(np.array(eps[0].views), want["eb_views0"])


def _device_vs_host_replay(device):
    """DeviceMemoryGroup == MemoryGroup row for row when it is given the same agent flush order and sample indices."""
    import torch
    from mfmarl_b200.algo import tools
    from mfmarl_b200.algo.replay_device import DeviceMemoryGroup
    gen = _gen()
    obs_shape, feat_shape, act_n, cap, E = (3, 3, 2), (5,), 4, 12, 3
    for use_mean in (False, True):
        np.random.seed(3)
        host = tools.MemoryGroup(obs_shape, feat_shape, act_n, max_len=200, batch_size=32, sub_len=10, use_mean=use_mean)
        dev = DeviceMemoryGroup(obs_shape, feat_shape, act_n, max_len=200, batch_size=32, sub_len=10, use_mean=use_mean,
                                device=device, stage_rows=E * cap * 10, id_span=1000)
        for rnd in range(3):
            streams = [list(gen.synthetic_stream(500 + 10 * rnd + e, 9 + e, 10, obs_shape, feat_shape, act_n)) for e in range(E)]
            T = max(len(s) for s in streams)
            for t in range(T):
                view = np.zeros((E, cap) + obs_shape, np.float32); feat = np.zeros((E, cap) + feat_shape, np.float32)
                acts = np.zeros((E, cap), np.int32); rew = np.zeros((E, cap), np.float32)
                alive = np.zeros((E, cap), np.uint8); ids = np.zeros((E, cap), np.int32)
                prob = np.zeros((E, act_n), np.float32); num = np.zeros((E,), np.int32); active = np.zeros((E,), bool)
                for e in range(E):
                    if t >= len(streams[e]):
                        continue
                    kw = streams[e][t]
                    n = len(kw["ids"])
                    view[e, :n], feat[e, :n], acts[e, :n], rew[e, :n] = kw["state"][0], kw["state"][1], kw["acts"], kw["rewards"]
                    alive[e, :n], ids[e, :n], prob[e], num[e], active[e] = kw["alives"], kw["ids"], kw["prob"][0], n, True
                    host.push(state=kw["state"], acts=kw["acts"], rewards=kw["rewards"], alives=kw["alives"],
                              ids=[e * 1000 + int(i) for i in kw["ids"]], prob=kw["prob"])
                tt = lambda a: torch.from_numpy(a).to(device)
                dev.push(state=(tt(view), tt(feat)), acts=tt(acts), rewards=tt(rew), alives=tt(alive), ids=tt(ids),
                         prob=tt(prob), num=tt(num), active=tt(active))
            # the host shuffles dict keys in insertion order with np.random.shuffle: replay that order on the device
            state = np.random.get_state()
            order = list(host.agent.keys())
            np.random.shuffle(order)
            np.random.set_state(state)
            host.tight()
            dev.tight(agent_order=order)
            assert dev.nb_entries == host.nb_entries
            n = host.nb_entries
            for a, b in ((dev.obs0, host.obs0), (dev.feat0, host.feat0), (dev.actions, host.actions),
                         (dev.rewards, host.rewards), (dev.terminals, host.terminals), (dev.masks, host.masks)):
                assert np.array_equal(a[:n].cpu().numpy(), b.pull()), rnd
            if use_mean:
                assert np.array_equal(dev.prob[:n].cpu().numpy(), host.prob.pull())
            assert dev.get_batch_num(False) == host.get_batch_num(False)
        state = np.random.get_state()
        idx = np.random.choice(host.nb_entries, size=host.batch_size)
        np.random.set_state(state)
        for a, b in zip(dev.sample(idx=idx), host.sample()):
            assert np.array_equal(a.cpu().numpy(), np.asarray(b))


def test_device_replay_matches_host_replay_cpu_tensors():
    _device_vs_host_replay("cpu")


@pytest.mark.gpu
def test_device_replay_matches_host_replay_cuda():
    _device_vs_host_replay("cuda")
