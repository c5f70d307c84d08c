"""GPU parity tests for the batched, device-resident path (mfb_* C ABI through mfmarl_b200):
every environment of a batch is compared bit for bit with its own C-oracle instance."""
import numpy as np
import pytest

from engines import OracleEngine
from lockstep import assert_same
from scenarios import c4_positions, fight_actions, generate_map_positions, uniform_actions

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def make(E, map_size=40, cap=64, pos=None, **kw):
    from mfmarl_b200 import BatchedGridWorld
    env = BatchedGridWorld(E, map_size=map_size, capacity=cap, **kw)
    env.reset()
    left, right = pos if pos is not None else generate_map_positions(map_size)
    env.add_agents(0, left)
    env.add_agents(1, right)
    oracles = []
    for _ in range(E):
        o = OracleEngine(map_size)
        o.reset()
        o.add_agents(0, left)
        o.add_agents(1, right)
        oracles.append(o)
    return env, oracles


def lockstep_batched(env, oracles, steps, seed, stream="fight", check_obs_every=1, min_deaths=0):
    E, cap = env.n_envs, env.capacity
    rngs = [np.random.RandomState(seed * 1000 + e) for e in range(E)]
    deaths = 0
    for s in range(steps):
        num = env.get_num()
        for e, o in enumerate(oracles):
            assert [o.get_num(0), o.get_num(1)] == list(num[e]), "num differs env %d step %d" % (e, s)
        if num.min() == 0:
            break
        if check_obs_every and s % check_obs_every == 0:
            view, feat = env.observe()
            view, feat = view.cpu().numpy(), feat.cpu().numpy()
            for e, o in enumerate(oracles):
                for g in range(2):
                    v, f = o.get_observation(g)
                    assert_same("view[e%d g%d]" % (e, g), v, view[e, g, :num[e, g]], s)
                    assert_same("feat[e%d g%d]" % (e, g), f, feat[e, g, :num[e, g]], s)
        pos = env.get("pos")
        actions = np.zeros((E, 2, cap), np.int32)
        for e, o in enumerate(oracles):
            for g in range(2):
                n = num[e, g]
                assert_same("pos[e%d g%d]" % (e, g), o.get_pos(g), pos[e, g, :n], s)
                a = fight_actions(rngs[e], pos[e, g, :n], env.map_size) if stream == "fight" \
                    else uniform_actions(rngs[e], n)
                actions[e, g, :n] = a
                o.set_action(g, a)
        reward, alive, done, mean = env.step(torch.from_numpy(actions).cuda())
        reward, alive, done, mean = reward.cpu().numpy(), alive.cpu().numpy(), done.cpu().numpy(), mean.cpu().numpy()
        for e, o in enumerate(oracles):
            d = o.step()
            assert bool(done[e]) == d, "done differs env %d step %d" % (e, s)
            for g in range(2):
                n = num[e, g]
                assert_same("reward[e%d g%d]" % (e, g), o.get_reward(g), reward[e, g, :n], s)
                al = o.get_alive(g)
                assert_same("alive[e%d g%d]" % (e, g), al, alive[e, g, :n].astype(bool), s)
                deaths += int((~al).sum())
                ref_mean = o.mean_action(actions[e, g, :n])
                np.testing.assert_allclose(mean[e, g], ref_mean, rtol=1e-6, atol=0)   # north_star tolerance
            o.clear_dead()
        hp = env.get("hp")
        num2 = env.get_num()
        for e, o in enumerate(oracles):
            for g in range(2):
                assert_same("hp[e%d g%d]" % (e, g), o.get_hp(g), hp[e, g, :num2[e, g]], s)
        if done.any():
            break
    assert deaths >= min_deaths, deaths
    return deaths


def test_batched_envs_with_independent_action_streams():
    env, oracles = make(6)
    lockstep_batched(env, oracles, steps=150, seed=1, min_deaths=50)


def test_batched_uniform_stream():
    env, oracles = make(3)
    lockstep_batched(env, oracles, steps=40, seed=2, stream="uniform")


def test_batched_80x80_512v512():
    env, oracles = make(2, map_size=80, cap=512, pos=c4_positions())
    lockstep_batched(env, oracles, steps=40, seed=3, check_obs_every=8, min_deaths=20)


def test_batched_ragged_capacity():
    """capacity not a multiple of the store chunk; n < capacity from the start"""
    left, right = generate_map_positions(40)
    env, oracles = make(2, cap=36, pos=(left[:29], right[:33]))
    assert env.capacity == 36
    lockstep_batched(env, oracles, steps=100, seed=4, min_deaths=1)


def test_mean_action_kernel_matches_numpy():
    from mfmarl_b200 import mean_action
    rng = np.random.RandomState(0)
    rows, cap = 37, 100
    acts = rng.randint(0, 21, size=(rows, cap)).astype(np.int32)
    num = rng.randint(1, cap + 1, size=rows).astype(np.int32)
    num[0] = cap
    out = mean_action(torch.from_numpy(acts).cuda(), torch.from_numpy(num).cuda()).cpu().numpy()
    for r in range(rows):
        # senario_battle.py:141 -- np.mean(one_hot(acts)) in float64, fed to the nets as fp32
        ref = np.mean(np.eye(21)[acts[r, :num[r]]], axis=0)
        np.testing.assert_allclose(out[r], ref, rtol=1e-6, atol=0)
        assert abs(out[r].sum() - 1.0) < 1e-5


def test_philox_results_do_not_depend_on_sharding():
    """Philox is keyed by the global env id: a shard that owns envs [2, 4) reproduces envs 2, 3 of the
    4-env engine exactly (multi-GPU sharding)."""
    from mfmarl_b200 import BatchedGridWorld
    left, right = generate_map_positions(40)

    def run(E, base):
        env = BatchedGridWorld(E, rng="philox", seed=7, env_base=base)
        env.reset(); env.add_agents(0, left); env.add_agents(1, right)
        outs = []
        for s in range(60):
            pos, num = env.get("pos"), env.get_num()
            acts = np.zeros((E, 2, 64), np.int32)
            for e in range(E):
                r = np.random.RandomState(100 * (base + e) + s)
                for g in range(2):
                    acts[e, g, :num[e, g]] = fight_actions(r, pos[e, g, :num[e, g]], 40)
            reward, alive, done, _ = env.step(torch.from_numpy(acts).cuda())
            outs.append((reward.cpu().numpy().copy(), alive.cpu().numpy().copy(), env.get("pos"), env.get("hp")))
        return outs

    full, shard = run(4, 0), run(2, 2)
    deaths = 0
    for (r4, a4, p4, h4), (r2, a2, p2, h2) in zip(full, shard):
        assert_same("reward", r4[2:], r2, 0); assert_same("alive", a4[2:], a2, 0)
        assert_same("pos", p4[2:], p2, 0); assert_same("hp", h4[2:], h2, 0)
        deaths += int((a4 == 0).sum())
    assert deaths > 0


def test_auto_reset_replaces_the_armies():
    from mfmarl_b200 import BatchedGridWorld
    left, right = generate_map_positions(40)
    env = BatchedGridWorld(2, max_steps=5, auto_reset=True)
    env.reset(); env.add_agents(0, left); env.add_agents(1, right)
    pos0, id0 = env.get("pos").copy(), env.get("id").copy()
    rng = np.random.RandomState(0)
    for s in range(5):
        acts = torch.from_numpy(rng.randint(0, 13, size=(2, 2, 64)).astype(np.int32)).cuda()
        env.step(acts)
        if s < 4:
            assert (env.get("step_ct") == s + 1).all()
            assert not np.array_equal(env.get("pos"), pos0)
    assert (env.get("step_ct") == 0).all()
    assert np.array_equal(env.get("pos"), pos0) and np.array_equal(env.get("id"), id0)
    assert (env.get("hp") == 10).all() and (env.get_num() == 64).all()
    assert (env.get("agent_steps") == 5 * 128).all()


def test_full_size_properties_4096_envs():
    """BASELINE config 3 at full size (4096 envs x 64 v 64): size-independent properties.
    (1) every env given the same actions equals env 0 (checksum of checksums), and env 0 equals the oracle;
    (2) one agent per cell, hp <= 10, alive bookkeeping; (3) observation identities per agent row."""
    E = 4096
    env, oracles = make(E)
    oracles = oracles[:1]
    o = oracles[0]
    rng = np.random.RandomState(5)
    for s in range(12):
        num = env.get_num()
        assert (num == num[0]).all()
        pos = env.get("pos")
        acts1 = np.zeros((1, 2, 64), np.int32)
        for g in range(2):
            acts1[0, g, :num[0, g]] = fight_actions(rng, pos[0, g, :num[0, g]], 40)
            o.set_action(g, acts1[0, g, :num[0, g]])
        view, feat = env.observe()
        if s % 4 == 0:
            # (3) own-has at the view centre, minimap channels sum to 1 (+1 self marker), id bits
            n0 = int(num[0, 0])
            v = view[:, 0, :n0]
            assert bool((v[:, :, 6, 6, 1] == 1).all())
            assert torch.allclose(v[..., 3].sum(dim=(-1, -2)), torch.full_like(v[:, :, 0, 0, 0], 2.0), atol=1e-5)
            assert torch.allclose(v[..., 6].sum(dim=(-1, -2)), torch.full_like(v[:, :, 0, 0, 0], 2.0), atol=1e-5)
            # (1) all envs identical to env 0
            assert bool((view[:, :, :n0] == view[0:1, :, :n0]).all())
            assert bool((feat[:, :, :n0] == feat[0:1, :, :n0]).all())
        actions = torch.from_numpy(acts1).cuda().expand(E, 2, 64).contiguous()
        reward, alive, done, mean = env.step(actions)
        d = o.step()
        assert bool((done == int(d)).all())
        for g in range(2):
            n = num[0, g]
            assert_same("reward", o.get_reward(g), reward[0, g, :n].cpu().numpy(), s)
            assert bool((reward[:, g, :n] == reward[0:1, g, :n]).all())
            assert bool((alive[:, g, :n] == alive[0:1, g, :n]).all())
        o.clear_dead()
        # (2) invariants on the whole batch
        pos, hp, num2 = env.get("pos"), env.get("hp"), env.get_num()
        assert hp.max() <= 10.0
        assert (num2 == [o.get_num(0), o.get_num(1)]).all()
        cells = pos[..., 1] * 40 + pos[..., 0]
        for e in (0, 1, E // 2, E - 1):
            live = np.concatenate([cells[e, g, :num2[e, g]] for g in range(2)])
            assert len(np.unique(live)) == len(live), "two agents in one cell"


def test_host_buffer_paths_match_the_device_path():
    """mfb_step_host (synchronous) and mfb_step_host_async (pipelined, two staging sets) give the results of
    mfb_step bit for bit (rows beyond num are unspecified and not compared)."""
    from mfmarl_b200 import BatchedGridWorld
    left, right = generate_map_positions(40)
    E = 5
    envs = []
    for _ in range(3):
        env = BatchedGridWorld(E, rng="philox", seed=3)
        env.reset(); env.add_agents(0, left); env.add_agents(1, right)
        envs.append(env)
    dev, sync_env, async_env = envs

    def pinned_set():
        return (torch.empty((E, 2, 64), dtype=torch.float32).pin_memory(),
                torch.empty((E, 2, 64), dtype=torch.uint8).pin_memory(),
                torch.empty((E, 2, 21), dtype=torch.float32).pin_memory(),
                torch.empty((E,), dtype=torch.int32).pin_memory())

    def same(got, want, num):
        for e in range(E):
            for g in range(2):
                n = num[e, g]
                assert torch.equal(got[0][e, g, :n], want[0][e, g, :n])
                assert torch.equal(got[1][e, g, :n], want[1][e, g, :n])
        assert torch.equal(got[2], want[2]) and torch.equal(got[3], want[3])

    sync_out, async_out = pinned_set(), [pinned_set(), pinned_set()]
    rng = np.random.RandomState(12)
    prev, deaths, ticket = None, 0, 0
    for s in range(80):
        pos, num = dev.get("pos"), dev.get_num()
        acts = np.zeros((E, 2, 64), np.int32)
        for e in range(E):
            for g in range(2):
                acts[e, g, :num[e, g]] = fight_actions(rng, pos[e, g, :num[e, g]], 40)
        h_act = torch.from_numpy(acts).pin_memory()
        r, a, d, m = dev.step(h_act.cuda())
        cur = tuple(t.cpu().clone() for t in (r, a, m, d))
        sync_env.step_host(h_act, *sync_out)
        same(sync_out, cur, num)
        ticket = async_env.step_host_async(h_act, *async_out[s & 1])
        if prev is not None:             # consume the previous step while this one is in flight
            async_env.host_wait(ticket ^ 1)
            same(async_out[(s - 1) & 1], prev[0], prev[1])
        prev = (cur, num.copy())
        for e in range(E):
            for g in range(2):
                deaths += int((cur[1][e, g, :num[e, g]] == 0).sum())
    async_env.host_wait(ticket)
    same(async_out[79 & 1], prev[0], prev[1])
    assert deaths > 10


def test_annihilated_group_keeps_stepping_without_nan():
    """A finished environment (one group wiped out) is unreachable in the reference's loop -- `done` ends it, and
    get_observation would dereference agents[0] (GridWorld.cc:357).  The lock-step engine keeps stepping it next to
    the live ones, so the behaviour is defined: done stays set, the empty group has no rows, its minimap and mean
    action are 0 (not 0/0), and the survivors' observations stay finite."""
    from mfmarl_b200 import BatchedGridWorld
    E = 3
    env = BatchedGridWorld(E, map_size=40, capacity=64, rng="minstd")
    env.reset()
    env.add_agents(0, [[10, 10, 0], [12, 10, 0], [10, 12, 0], [12, 12, 0]])      # four attackers around ...
    env.add_agents(1, [[11, 11, 0]])                                              # ... one victim
    # attack deltas: 13 (-1,-1) 15 (1,-1) 18 (-1,1) 20 (1,1); the victim idles
    acts = torch.zeros((E, 2, 64), dtype=torch.int32, device="cuda")
    acts[:, 0, :4] = torch.tensor([20, 18, 15, 13], dtype=torch.int32)
    acts[:, 1, 0] = 6
    done_at = None
    for s in range(6):
        env.observe()
        reward, alive, done, mean = env.step(acts)
        if int(done[0]) and done_at is None:
            done_at = s
    assert done_at is not None and done_at <= 2        # 4 hits of 2 per step against hp 10 (+0.1 recovery): dead in 2 steps
    assert env.get_num().tolist() == [[4, 0]] * E
    view, feat = env.observe()
    assert torch.isfinite(view[:, 0, :4]).all() and torch.isfinite(feat[:, 0, :4]).all()
    assert float(view[:, 0, :4, :, :, 6].abs().max()) <= 1.0           # other-group minimap: only the +1 self marker
    reward, alive, done, mean = env.step(acts)
    assert done.tolist() == [1] * E and torch.isfinite(reward[:, 0, :4]).all()
    assert float(mean[:, 1].abs().sum()) == 0.0 and abs(float(mean[0, 0].sum()) - 1.0) < 1e-6


def test_batched_largest_supported_shape_128x128_1024v1024():
    """Maximum sizes: 1024 agent slots per group is the widest k_step CTA (1024 threads) and a 128x128 map the
    largest occupancy grid exercised; dense placement (stride 1 blocks, one empty column apart) so that attacks, kills and
    move collisions all happen.  Every output bit against the C oracle."""
    from scenarios import block_positions
    pos = (block_positions(48, 32, 16, 64, stride=1), block_positions(65, 32, 16, 64, stride=1))
    assert len(pos[0]) == len(pos[1]) == 1024
    env, oracles = make(2, map_size=128, cap=1024, pos=pos)
    assert env.capacity == 1024
    lockstep_batched(env, oracles, steps=24, seed=9, check_obs_every=6, min_deaths=10)


@pytest.mark.parametrize("seed", range(8))
def test_randomized_shapes_walls_and_ragged_armies(seed):
    """Differential fuzzing: random map side (odd and even, small and large), random interior walls, armies of
    random ragged sizes scattered at random free cells, three environments with independent action streams --
    every output bit against the C oracle."""
    from mfmarl_b200 import BatchedGridWorld
    rng = np.random.RandomState(100 + seed)
    size = int(rng.choice([14, 17, 23, 32, 40, 57, 64, 90]))
    cells = [(x, y) for x in range(1, size - 1) for y in range(1, size - 1)]
    rng.shuffle(cells)
    n_wall = int(rng.randint(0, size))                       # a few interior obstacles
    n0 = int(rng.randint(3, min(150, len(cells) // 6)))
    n1 = int(rng.randint(3, min(150, len(cells) // 6)))
    walls, rest = cells[:n_wall], cells[n_wall:]
    # keep the armies close so that they fight: the n0 + n1 free cells nearest to the map centre
    rest.sort(key=lambda c: (c[0] - size // 2) ** 2 + (c[1] - size // 2) ** 2)
    picked = rest[:n0 + n1]
    rng.shuffle(picked)
    g0 = [[x, y, 0] for x, y in picked[:n0]]
    g1 = [[x, y, 0] for x, y in picked[n0:]]
    cap = int(rng.choice([max(n0, n1), max(n0, n1) + 3, 160]))
    E = 3
    env = BatchedGridWorld(E, map_size=size, capacity=cap, rng="minstd")
    env.reset()
    if walls:
        env.add_walls([[x, y, 0] for x, y in walls])
    env.add_agents(0, g0); env.add_agents(1, g1)
    oracles = []
    for _ in range(E):
        o = OracleEngine(size)
        o.reset()
        if walls:
            o.add_walls([[x, y, 0] for x, y in walls])
        o.add_agents(0, g0); o.add_agents(1, g1)
        oracles.append(o)
    lockstep_batched(env, oracles, steps=45, seed=200 + seed, check_obs_every=3)


def test_sampled_envs_of_a_large_engine_match_the_oracle():
    """512 lock-stepped environments, three of them (first, middle, last) shadowed by the C oracle on the fight stream
    while the others run uniform random actions: indexing at scale (env offsets, ticket scheduling, per-env rng)."""
    from mfmarl_b200 import BatchedGridWorld
    E, cap, steps = 512, 64, 150
    watch = [0, 255, 511]
    left, right = generate_map_positions(40)
    env = BatchedGridWorld(E, map_size=40, capacity=cap, rng="minstd")
    env.reset(); env.add_agents(0, left); env.add_agents(1, right)
    oracles = {}
    for e in watch:
        o = OracleEngine(40); o.reset(); o.add_agents(0, left); o.add_agents(1, right)
        oracles[e] = o
    rng = np.random.RandomState(77)
    gen = torch.Generator(device="cuda"); gen.manual_seed(5)
    deaths = 0
    for s in range(steps):
        num, pos = env.get_num(), env.get("pos")
        view, feat = env.observe()
        actions = torch.randint(0, 21, (E, 2, cap), generator=gen, device="cuda", dtype=torch.int32)
        for e, o in oracles.items():
            for g in range(2):
                n = num[e, g]
                assert o.get_num(g) == n
                if s % 10 == 0:
                    v, f = o.get_observation(g)
                    assert_same("view[e%d g%d]" % (e, g), v, view[e, g, :n].cpu().numpy(), s)
                    assert_same("feat[e%d g%d]" % (e, g), f, feat[e, g, :n].cpu().numpy(), s)
                a = fight_actions(rng, pos[e, g, :n], 40)
                actions[e, g, :n] = torch.from_numpy(a).cuda()
                o.set_action(g, a)
        reward, alive, done, mean = env.step(actions)
        reward, alive = reward.cpu().numpy(), alive.cpu().numpy()
        for e, o in oracles.items():
            assert o.step() == bool(done[e])
            for g in range(2):
                n = num[e, g]
                assert_same("reward[e%d g%d]" % (e, g), o.get_reward(g), reward[e, g, :n], s)
                al = o.get_alive(g)
                assert_same("alive[e%d g%d]" % (e, g), al, alive[e, g, :n].astype(bool), s)
                deaths += int((~al).sum())
            o.clear_dead()
        if any(min(o.get_num(0), o.get_num(1)) == 0 for o in oracles.values()):
            break
    assert deaths > 30


def per_env_placements(E, map_size, rng):
    """what generate_map does per round (senario_battle.py:8-37): which army takes the left block is random per env
    (the block added FIRST gets the low ids); plus per-env noise -- a few duplicated, walled and out-of-range cells
    that the engine must skip (GridWorld.cc:180-187), so the envs hold different numbers of agents"""
    left, right = generate_map_positions(map_size)
    pos = [[], []]
    for e in range(E):
        swap = int(rng.randint(2))
        blocks = [left, right] if not swap else [right, left]           # blocks[g] = where group g starts
        for g in range(2):
            p = blocks[g][:, :2].copy()
            k = rng.randint(0, 6)                                        # overwrite a few entries with bad cells
            for j in rng.choice(len(p), size=k, replace=False):
                p[j] = [(0, 7), (map_size - 1, 3), tuple(p[(j + 1) % len(p)]), (-2, 5)][rng.randint(4)]
            pos[g].append(p)
    return np.stack(pos[0]), np.stack(pos[1])


def test_per_env_placements_match_independent_oracles():
    """mfb_add_agents_per_env: every env its own armies.  Each env is compared with its own oracle
    that was given the same positions in the same order."""
    from mfmarl_b200 import BatchedGridWorld
    E = 6
    rng = np.random.RandomState(44)
    p0, p1 = per_env_placements(E, 40, rng)
    env = BatchedGridWorld(E, map_size=40, capacity=64)
    env.reset()
    oracles = []
    added0 = env.add_agents_per_env(0, p0)
    added1 = env.add_agents_per_env(1, p1)
    for e in range(E):
        o = OracleEngine(40)
        o.reset()
        assert o.add_agents(0, np.c_[p0[e], np.zeros(len(p0[e]), np.int32)]) == added0[e]
        assert o.add_agents(1, np.c_[p1[e], np.zeros(len(p1[e]), np.int32)]) == added1[e]
        oracles.append(o)
    num = env.get_num()
    assert len({tuple(n) for n in num}) > 1, "the placements were meant to leave different counts: %s" % num
    for e, o in enumerate(oracles):
        assert [o.get_num(0), o.get_num(1)] == list(num[e])
        for g in range(2):
            assert_same("id", o.get_agent_id(g), env.get("id")[e, g, :num[e, g]], 0)
    lockstep_batched(env, oracles, steps=40, seed=9, stream="fight", min_deaths=10)


def test_shared_and_per_env_adds_mix_and_auto_reset_uses_each_envs_template():
    from mfmarl_b200 import BatchedGridWorld
    E = 4
    left, right = generate_map_positions(40)
    env = BatchedGridWorld(E, map_size=40, capacity=64, max_steps=3, auto_reset=True)
    env.reset()
    env.add_agents(0, left)                                   # shared ...
    extra = np.array([[[20, 3 + e], [20, 3 + e], [21, 36 - e]] for e in range(E)], np.int32)
    assert list(env.add_agents_per_env(1, extra)) == [2] * E  # ... then per env (one duplicate skipped)
    env.add_agents(1, right[:10])                             # ... and shared again: lands in every env's own list
    pos0, id0, num0 = env.get("pos").copy(), env.get("id").copy(), env.get_num().copy()
    assert (num0 == [64, 12]).all()
    for e in range(E):
        assert pos0[e, 1, 0].tolist() == [20, 3 + e] and pos0[e, 1, 1].tolist() == [21, 36 - e]
        assert id0[e, 1, :12].tolist() == list(range(64, 76))
    rng = np.random.RandomState(1)
    for s in range(3):
        env.step(torch.from_numpy(rng.randint(0, 13, size=(E, 2, 64)).astype(np.int32)).cuda())
    assert (env.get("step_ct") == 0).all() and (env.get("episode") == 1).all()
    assert np.array_equal(env.get("pos"), pos0) and np.array_equal(env.get("id"), id0) and np.array_equal(env.get_num(), num0)


def test_random_sides_at_auto_reset():
    """random_sides: at every auto-reset an env draws whether the armies swap blocks -- positions AND ids, since
    generate_map adds the left army first (senario_battle.py:14-37).  Keyed by (seed, global env id, episode): the
    draws do not depend on how the envs are sharded over engines."""
    from mfmarl_b200 import BatchedGridWorld
    E = 64
    left, right = generate_map_positions(40)

    def run(n, base):
        env = BatchedGridWorld(n, map_size=40, capacity=64, max_steps=2, auto_reset=True, random_sides=True,
                               rng="philox", seed=11, env_base=base)
        env.reset(); env.add_agents(0, left); env.add_agents(1, right)
        sides = []
        acts = torch.zeros((n, 2, 64), dtype=torch.int32, device="cuda") + 6        # everybody stays put
        for ep in range(6):
            env.step(acts); env.step(acts)
            side, pos, ids = env.get("side"), env.get("pos"), env.get("id")
            assert (env.get("episode") == ep + 1).all()
            for e in range(n):
                a, b = (right, left) if side[e] else (left, right)
                assert np.array_equal(pos[e, 0, :64], a[:, :2]) and np.array_equal(pos[e, 1, :64], b[:, :2])
                lo, hi = np.arange(64), np.arange(64, 128)
                assert np.array_equal(ids[e, 0, :64], hi if side[e] else lo)
                assert np.array_equal(ids[e, 1, :64], lo if side[e] else hi)
            sides.append(side.copy())
        return np.stack(sides)

    full = run(E, 0)
    assert 0.3 < full.mean() < 0.7 and len({tuple(r) for r in full}) > 1
    assert np.array_equal(run(16, 32), full[:, 32:48])


@pytest.mark.parametrize("map_size,cap,E", [(40, 64, 5), (80, 512, 2)])
def test_bf16_observation_rows_are_the_fp32_rows_rounded(map_size, cap, E):
    """mfb_observe_groups_bf16: [E, cap, 13, 13, 8] bf16 rows = the fp32 observation rounded to nearest-even bf16 with a
    zero eighth channel -- bit for bit, mid-fight (dead agents gone, hp fractions, minimap fractions, self markers), for
    both set-up modes of k_obs (rebuild at cap 64, observation record at cap 512)."""
    pos = generate_map_positions(40) if map_size == 40 else c4_positions()
    env, oracles = make(E, map_size=map_size, cap=cap, pos=pos)
    lockstep_batched(env, oracles, steps=12, seed=3, stream="fight", check_obs_every=0)
    for rnd in range(3):
        num = env.get_num()
        f32 = env.observe_groups()
        b16 = env.observe_groups(dtype=torch.bfloat16)
        for g in range(2):
            view32, feat32 = f32[g]
            view16, feat16 = b16[g]
            assert view16.dtype == torch.bfloat16 and tuple(view16.shape) == (E, cap, 13, 13, 8)
            for e in range(E):
                n = int(num[e, g])
                want = torch.zeros((n, 13, 13, 8), dtype=torch.bfloat16, device=view32.device)
                want[..., :7] = view32[e, :n].to(torch.bfloat16)
                assert torch.equal(view16[e, :n].view(torch.int16), want.view(torch.int16)), (rnd, g, e)
                assert torch.equal(feat16[e, :n], feat32[e, :n])
                ov, _ = oracles[e].get_observation(g)              # and the fp32 rows are the oracle's
                assert np.array_equal(ov.view(np.uint32), view32[e, :n].cpu().numpy().view(np.uint32))
        lockstep_batched(env, oracles, steps=4, seed=10 + rnd, stream="fight", check_obs_every=0)


def test_observe_leaves_slots_for_a_pipelined_sibling_and_gives_the_same_rows():
    """concurrent_step_envs only changes how many persistent k_obs CTAs are launched (room for the sibling engine's
    k_step on another stream); the observations are the same."""
    from mfmarl_b200 import BatchedGridWorld
    left, right = c4_positions()
    envs = []
    for reserve in (0, 4):
        env = BatchedGridWorld(4, map_size=80, capacity=512, concurrent_step_envs=reserve)
        env.reset(); env.add_agents(0, left); env.add_agents(1, right)
        envs.append(env)
    rng = np.random.RandomState(2)
    for s in range(3):
        acts = torch.from_numpy(rng.randint(0, 21, size=(4, 2, 512)).astype(np.int32)).cuda()
        views = []
        for env in envs:
            v, f = env.observe()
            views.append((v.clone(), f.clone()))
            env.step(acts)
        assert torch.equal(views[0][0], views[1][0]) and torch.equal(views[0][1], views[1][1])


@pytest.mark.parametrize("map_size,cap", [(40, 64), (80, 512)])
def test_convoys_and_focused_fire_resolve_like_the_ordered_loops(map_size, cap):
    """Worst cases for k_step's parallel formulations: packed blocks (stride 1) in which EVERYBODY makes the same move,
    so that almost every mover's target is held by another mover whose own fate decides (long chains for the pointer
    jumping, first-come-first-served ties on every freed cell), alternating with steps in which everybody attacks the
    same way (many attacks per victim, attackers that are victims: the entangled path), and a few random steps."""
    from scenarios import block_positions
    n_side = int(np.sqrt(cap))
    rows = cap // n_side
    left = block_positions(2, 3, n_side, rows, stride=1)
    right = block_positions(2 + n_side, 3, n_side, rows, stride=1)        # the two armies touch
    env, oracles = make(2, map_size=map_size, cap=cap, pos=(left, right))
    E = 2
    rng = np.random.RandomState(map_size)
    plan = [7, 7, 10, 5, 16, 17, 2, 8, 4, 14, 19, 7, 11, 1, 16, 6, 12, 0]        # moves (0-12) and attacks (13-20), see SURVEY 8
    deaths = 0
    for s in range(3 * len(plan)):
        num = env.get_num()
        pos = env.get("pos")
        actions = np.zeros((E, 2, cap), np.int32)
        for e, o in enumerate(oracles):
            for g in range(2):
                n = num[e, g]
                assert o.get_num(g) == n
                assert_same("pos", o.get_pos(g), pos[e, g, :n], s)
                a = np.full(n, plan[s % len(plan)], np.int32)
                if s % 5 == 4:
                    a = uniform_actions(rng, n)
                elif e == 1:                       # env 1: the two groups do different things
                    a[:] = plan[(s + 3 * g) % len(plan)]
                actions[e, g, :n] = a
                o.set_action(g, a)
        reward, alive, done, mean = env.step(torch.from_numpy(actions).cuda())
        reward, alive = reward.cpu().numpy(), alive.cpu().numpy()
        for e, o in enumerate(oracles):
            o.step()
            for g in range(2):
                n = num[e, g]
                assert_same("reward", o.get_reward(g), reward[e, g, :n], s)
                al = o.get_alive(g)
                assert_same("alive", al, alive[e, g, :n].astype(bool), s)
                deaths += int((~al).sum())
            o.clear_dead()
        if s % 6 == 0:
            view, feat = env.observe()
            view = view.cpu().numpy()
            num2 = env.get_num()
            for e, o in enumerate(oracles):
                for g in range(2):
                    v, f = o.get_observation(g)
                    assert_same("view", v, view[e, g, :num2[e, g]], s)
        if (env.get_num() == 0).any():
            break
    assert deaths > 0, deaths


def test_philox_attack_shuffle_uses_the_documented_keys():
    """rng="philox": the attack shuffle of env e at its step s is the reference's inside-out Fisher-Yates
    (GridWorld.cc:510-515) with draw i = Philox4x32-10(counter (i, s, 0, 0), key (seed, env_base + e)).x % (i + 1)
    (battle_kernels.cuh).  The oracle gets exactly that permutation injected, computed in numpy from the published
    round function (tests/philox_ref.py), and every output must then agree bit for bit."""
    from philox_ref import philox4x32_10
    E, seed, base = 3, 9, 5
    env, oracles = make(E, rng="philox", seed=seed, env_base=base)
    rngs = [np.random.RandomState(40 + e) for e in range(E)]
    deaths = 0
    for s in range(80):
        num, pos = env.get_num(), env.get("pos")
        actions = np.zeros((E, 2, env.capacity), np.int32)
        for e, o in enumerate(oracles):
            n_att = 0
            for g in range(2):
                n = num[e, g]
                a = fight_actions(rngs[e], pos[e, g, :n], env.map_size)
                actions[e, g, :n] = a
                o.set_action(g, a)
                n_att += int((a >= 13).sum())
            i = np.arange(n_att, dtype=np.uint32)
            draws = philox4x32_10([i, np.full_like(i, s), np.zeros_like(i), np.zeros_like(i)],
                                  [np.full_like(i, seed), np.full_like(i, base + e)])[0] % (i + 1)
            order = np.arange(n_att, dtype=np.int32)
            for k in range(n_att):
                j = int(draws[k]); order[k], order[j] = order[j], order[k]
            o.inject_attack_order(order)
        reward, alive, done, _ = env.step(torch.from_numpy(actions).cuda())
        reward, alive = reward.cpu().numpy(), alive.cpu().numpy()
        for e, o in enumerate(oracles):
            assert o.step() == bool(done[e])
            for g in range(2):
                n = num[e, g]
                assert_same("reward[e%d g%d]" % (e, g), o.get_reward(g), reward[e, g, :n], s)
                al = o.get_alive(g)
                assert_same("alive[e%d g%d]" % (e, g), al, alive[e, g, :n].astype(bool), s)
                deaths += int((~al).sum())
            o.clear_dead()
        if done.any():
            break
    assert deaths > 10, deaths
