"""Run two engines in lockstep on one action stream and compare every output bit for bit."""
import numpy as np

from scenarios import fight_actions, uniform_actions


def bits(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:
        return a.view(np.uint32)
    return a


def assert_same(name, a, b, step):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, "%s shape %s vs %s at step %d" % (name, a.shape, b.shape, step)
    if not np.array_equal(bits(a), bits(b)):
        bad = np.argwhere(bits(a) != bits(b))
        raise AssertionError("%s differs at step %d: %d mismatches, first at %s: %r vs %r" % (
            name, step, len(bad), bad[0], a[tuple(bad[0])], b[tuple(bad[0])]))


def setup_pair(engines, pos0, pos1, walls=None):
    for e in engines:
        e.reset()
        if walls is not None and len(walls):
            e.add_walls(walls)
        e.add_agents(0, pos0)
        e.add_agents(1, pos1)


def run_lockstep(A, B, steps, seed, stream="fight", check_obs_every=1, skip_clear_every=0,
                 on_step=None):
    """A is the trusted engine (drives the action stream from its positions); B is under test.
    Returns stats.  Raises AssertionError on the first differing bit."""
    rng = np.random.RandomState(seed)
    deaths = 0
    agent_steps = 0
    for s in range(steps):
        n = [A.get_num(g) for g in range(2)]
        assert n == [B.get_num(g) for g in range(2)], "num differs at step %d" % s
        if min(n) == 0:
            break
        if check_obs_every and s % check_obs_every == 0:
            for g in range(2):
                va, fa = A.get_observation(g)
                vb, fb = B.get_observation(g)
                assert_same("view[g%d]" % g, va, vb, s)
                assert_same("feature[g%d]" % g, fa, fb, s)
        for g in range(2):
            assert_same("id[g%d]" % g, A.get_agent_id(g), B.get_agent_id(g), s)
            assert_same("pos[g%d]" % g, A.get_pos(g), B.get_pos(g), s)
        acts = []
        for g in range(2):
            if stream == "fight":
                a = fight_actions(rng, A.get_pos(g), A.map_size)
            else:
                a = uniform_actions(rng, n[g])
            acts.append(a)
        for g in range(2):
            A.set_action(g, acts[g])
            B.set_action(g, acts[g])
        if on_step is not None:
            on_step(s, A, B, acts)
        da, db = A.step(), B.step()
        assert da == db, "done differs at step %d" % s
        for g in range(2):
            assert_same("reward[g%d]" % g, A.get_reward(g), B.get_reward(g), s)
            al = A.get_alive(g)
            assert_same("alive[g%d]" % g, al, B.get_alive(g), s)
            assert_same("pos_after[g%d]" % g, A.get_pos(g), B.get_pos(g), s)
            deaths += int((~al).sum())
        agent_steps += sum(n)
        if not (skip_clear_every and s % skip_clear_every == skip_clear_every - 1):
            A.clear_dead()
            B.clear_dead()
        if da:
            break
    return {"steps": s + 1, "deaths": deaths, "agent_steps": agent_steps,
            "final_num": [A.get_num(g) for g in range(2)]}
