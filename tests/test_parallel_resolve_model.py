"""Host-side models of the parallel formulations k_step uses for the three ORDER-DEPENDENT sections of
GridWorld::step, each checked against the one-at-a-time loop the reference runs (CPU test, no GPU):

  shuffle   inside-out Fisher-Yates of the attack list (GridWorld.cc:510-515): position p of the result is found
            by tracing the LAST swap that touched it -- no swap is ever executed
  attacks   GridWorld.cc:524-557 + Map::do_attack (Map.cc:266-321): attacks that cannot interact with any other
            attack of the step ("isolated") are resolved all at once; only the entangled rest keeps its order
  moves     GridWorld.cc:631-672 + Map::do_move (Map.cc:324-369): first come first served per target cell is a
            minimum over mover indices past the occupant's own turn, and "did the occupant leave" is a chain
            that is followed to its head

The kernel code in csrc/battle_kernels.cuh is a transcription of the *_parallel functions below; these tests
are the exactness argument for it, run on thousands of random instances.
"""
import numpy as np
import pytest


# ----------------------------------------------------------------------------------------------- shuffle
def shuffle_sequential(j):
    a = list(range(len(j)))
    for i, ji in enumerate(j):
        a[i], a[ji] = a[ji], a[i]
    return a


def shuffle_parallel(j):
    """result[p] for every p independently.  Step i swaps positions i and j_i <= i, and position i still holds
    element i when step i runs.  The content of position q after all steps < t is therefore: element i* if i* is the
    last step in (q, t) with j_i* = q; else whatever position j_q held after the steps < q (step q moved it in)."""
    n = len(j)
    head = [-1] * n          # linked lists: steps i that swap INTO position q (j_i = q), any order
    nxt = [-1] * n
    for i in range(n):       # (the kernel: nxt[i] = atomicExch(&head[j_i], i))
        nxt[i] = head[j[i]]
        head[j[i]] = i
    out = []
    hops = 0
    for p in range(n):
        q, t = p, n
        while True:
            best = -1
            i = head[q]
            while i >= 0:
                if q < i < t and i > best:
                    best = i
                i = nxt[i]
            if best >= 0:
                out.append(best)
                break
            if j[q] == q:
                out.append(q)
                break
            q, t = j[q], q
            hops += 1
    return out, hops


@pytest.mark.parametrize("n", [0, 1, 2, 3, 31, 32, 33, 100, 390, 1024])
def test_shuffle_trace_equals_the_swap_loop(n):
    rng = np.random.RandomState(n)
    for _ in range(40 if n < 500 else 5):
        j = [int(rng.randint(0, i + 1)) for i in range(n)]
        par, hops = shuffle_parallel(j)
        assert par == shuffle_sequential(j)
        assert sorted(par) == list(range(n))


def test_shuffle_trace_degenerate_draws():
    for j in ([0] * 50, list(range(50)), [max(0, i - 1) for i in range(50)], [i // 2 for i in range(64)]):
        assert shuffle_parallel(j)[0] == shuffle_sequential(j)


# ----------------------------------------------------------------------------------------------- attacks
DAMAGE, KILL_REWARD, DEAD_PENALTY, ATTACK_PENALTY = 2.0, 5.0, -0.1, -0.1
OP_NULL, OP_ATTACK, OP_KILL = 11, 7, 3


def attack_one(k, v, hp, dead, nr, op, acted, i):
    """one iteration of GridWorld.cc:524-557 (kill_supply = 0)"""
    if dead[k]:
        return
    acted[i] = True
    if v < 0 or dead[v]:
        nr[k] = np.float32(nr[k] + np.float32(ATTACK_PENALTY))
        return
    hp[v] = np.float32(hp[v] - np.float32(DAMAGE))
    reward = np.float32(0.0)
    if hp[v] < 0:
        dead[v] = True
        nr[v] = np.float32(DEAD_PENALTY)
        op[k] = OP_KILL
        reward = np.float32(KILL_REWARD)
    else:
        op[k] = OP_ATTACK
    nr[k] = np.float32(nr[k] + np.float32(reward + np.float32(ATTACK_PENALTY)))


def attacks_sequential(att, hp, dead, nr, op):
    acted = [False] * len(att)
    for i, (k, v) in enumerate(att):
        attack_one(k, v, hp, dead, nr, op, acted, i)
    return acted


def attacks_parallel(att, hp, dead, nr, op, rng):
    """Isolated attack (k -> v): nobody attacks k this step (so k is alive at its turn) and either it hits nothing
    or it is the only attack on v AND it cannot change what v itself does (v does not attack, or survives the hit).
    Such an attack reads and writes nothing another attack reads or writes; all of them run first, in ANY order
    (here: a random one), then the entangled attacks run one by one in list order."""
    n_slots = len(hp)
    n_in = [0] * n_slots
    attacks_out = [False] * n_slots
    for k, v in att:
        attacks_out[k] = True
        if v >= 0:
            n_in[v] += 1
    isolated = []
    for k, v in att:
        iso = n_in[k] == 0 and (v < 0 or (n_in[v] == 1 and
                                           (not attacks_out[v] or not (np.float32(hp[v] - np.float32(DAMAGE)) < 0))))
        isolated.append(iso)
    acted = [False] * len(att)
    order = [i for i in range(len(att)) if isolated[i]]
    rng.shuffle(order)
    for i in order:
        attack_one(att[i][0], att[i][1], hp, dead, nr, op, acted, i)
    for i in range(len(att)):
        if not isolated[i]:
            attack_one(att[i][0], att[i][1], hp, dead, nr, op, acted, i)
    return acted, sum(isolated)


@pytest.mark.parametrize("n_slots,density", [(16, 0.9), (64, 0.5), (128, 0.3), (1024, 0.4), (40, 1.0)])
def test_attack_split_equals_the_ordered_loop(n_slots, density):
    rng = np.random.RandomState(n_slots)
    iso_total = 0
    for trial in range(60 if n_slots < 500 else 6):
        attackers = [k for k in range(n_slots) if rng.random_sample() < density]
        rng.shuffle(attackers)
        # victims: concentrated on a few slots so that chains (victim attacks back, several attackers per victim) are common
        hot = rng.randint(0, n_slots, size=max(2, n_slots // 6))
        att = []
        for k in attackers:
            r = rng.random_sample()
            v = -1 if r < 0.3 else int(hot[rng.randint(len(hot))]) if r < 0.8 else int(rng.randint(n_slots))
            if v == k:
                v = -1
            att.append((k, v))
        hp0 = rng.choice(np.array([0.5, 1.9, 2.0, 3.7, 4.0, 10.0], np.float32), size=n_slots)
        state = lambda: (hp0.copy(), [False] * n_slots, np.full(n_slots, -0.005, np.float32), [OP_NULL] * n_slots)
        a, b = state(), state()
        acted_a = attacks_sequential(att, *a)
        acted_b, n_iso = attacks_parallel(att, *b, rng)
        iso_total += n_iso
        assert acted_a == acted_b
        assert np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32))
        assert a[1] == b[1] and a[3] == b[3]
        assert np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32))
    assert iso_total > 0


# ----------------------------------------------------------------------------------------------- moves
EMPTY, WALL = 0, 1
OP_COLLIDE = 6


def moves_sequential(W, H, grid, pos, dead, movers, op):
    """movers: list of (slot, dx, dy) in processing order.  Map::do_move: out of the board -> nothing; target blank
    and free (or the mover's own cell) -> move; held by an agent -> collide; wall -> nothing."""
    for k, dx, dy in movers:
        if dead[k]:
            continue
        x, y = pos[k]
        nx, ny = x + dx, y + dy
        if nx < 0 or ny < 0 or nx + 1 >= W or ny + 1 >= H:
            continue
        occ = grid[ny][nx]
        if occ == EMPTY or occ == 2 + k:
            grid[y][x] = EMPTY
            grid[ny][nx] = 2 + k
            pos[k] = (nx, ny)
        elif occ >= 2:
            op[k] = OP_COLLIDE


def moves_parallel(W, H, grid, pos, dead, movers, op):
    n = len(movers)
    UNKNOWN = -1
    cand = [False] * n
    tgt = [None] * n
    leaves = {}                       # slot -> mover index, for movers that really try to leave their cell
    for m, (k, dx, dy) in enumerate(movers):
        if dead[k]:
            continue
        x, y = pos[k]
        nx, ny = x + dx, y + dy
        if nx < 0 or ny < 0 or nx + 1 >= W or ny + 1 >= H:
            continue
        cand[m], tgt[m] = True, (nx, ny)
        if (nx, ny) != (x, y):
            leaves[k] = m
    # phase 1: classify against the grid as it is BEFORE any move; contenders race for the cell with a min over m
    cellmin = {}
    res = [0] * n                      # 0 fails / nothing, 1 moves, UNKNOWN waits for its dependency
    dep = [-1] * n
    collide = [False] * n
    contender = [False] * n
    for m in range(n):
        if not cand[m]:
            continue
        k = movers[m][0]
        nx, ny = tgt[m]
        occ = grid[ny][nx]
        if occ == 2 + k:               # (0, 0): stays, trivially succeeds, blocks everybody else for the whole step
            continue
        if occ == WALL:
            continue                   # never free, no collide object
        if occ == EMPTY:
            thr, d = -1, -1
        else:
            o = occ - 2
            j = leaves.get(o, -1)
            if j < 0 or m < j:         # the occupant never leaves, or has not had its turn yet
                collide[m] = True
                continue
            thr, d = j, j
        contender[m], dep[m] = True, d
        key = (nx, ny)
        cellmin[key] = min(cellmin.get(key, n), m)
    # phase 2: the first contender inherits the cell iff it was empty or its occupant manages to leave
    for m in range(n):
        if contender[m]:
            if cellmin[tgt[m]] == m:
                res[m] = 1 if dep[m] < 0 else UNKNOWN
            else:
                collide[m] = True      # the cell is held at its turn: by the winner, or still by the occupant
    rounds = 0
    while any(r == UNKNOWN for r in res):
        rounds += 1
        new = list(res)
        for m in range(n):
            if res[m] == UNKNOWN and res[dep[m]] != UNKNOWN:
                new[m] = res[dep[m]]
                if new[m] == 0:
                    collide[m] = True
        res = new
    # phase 3: commit -- vacate, then enter
    for m in range(n):
        if res[m] == 1:
            x, y = pos[movers[m][0]]
            grid[y][x] = EMPTY
    for m in range(n):
        k = movers[m][0]
        if res[m] == 1:
            nx, ny = tgt[m]
            grid[ny][nx] = 2 + k
            pos[k] = (nx, ny)
        if collide[m]:
            op[k] = OP_COLLIDE
    return rounds


MOVES = [(0, -2), (-1, -1), (0, -1), (1, -1), (-2, 0), (-1, 0), (0, 0), (1, 0), (2, 0), (-1, 1), (0, 1), (1, 1), (0, 2)]


def random_world(rng, W, H, n_agents, p_move, p_dead, bias=None):
    grid = [[EMPTY] * W for _ in range(H)]
    for x in range(W):
        grid[0][x] = grid[H - 1][x] = WALL
    for y in range(H):
        grid[y][0] = grid[y][W - 1] = WALL
    cells = [(x, y) for y in range(1, H - 1) for x in range(1, W - 1)]
    rng.shuffle(cells)
    for x, y in cells[n_agents:n_agents + max(1, n_agents // 8)]:
        grid[y][x] = WALL              # a few inner walls
    pos, dead = [], []
    for k, (x, y) in enumerate(cells[:n_agents]):
        pos.append((x, y))
        d = rng.random_sample() < p_dead
        dead.append(d)
        if not d:
            grid[y][x] = 2 + k
    order = list(range(n_agents))
    rng.shuffle(order)
    movers = []
    for k in order:
        if rng.random_sample() < p_move:
            dx, dy = MOVES[rng.randint(13)] if bias is None or rng.random_sample() < 0.3 else bias
            movers.append((k, dx, dy))
    return grid, pos, dead, movers


@pytest.mark.parametrize("W,H,n_agents,p_move,bias", [
    (8, 8, 20, 0.9, None), (12, 9, 60, 1.0, None), (12, 12, 90, 1.0, (1, 0)), (40, 40, 128, 0.6, None),
    (20, 6, 70, 1.0, (-1, 0)), (16, 16, 190, 1.0, (0, 1)), (30, 30, 500, 0.8, (2, 0)), (80, 80, 1024, 0.62, None)])
def test_move_resolution_equals_the_ordered_loop(W, H, n_agents, p_move, bias):
    rng = np.random.RandomState(W * 1000 + n_agents)
    max_rounds = 0
    for trial in range(80 if n_agents < 400 else 8):
        grid, pos, dead, movers = random_world(rng, W, H, n_agents, p_move, 0.1, bias)
        ga, pa, oa = [r[:] for r in grid], pos[:], [OP_NULL] * n_agents
        gb, pb, ob = [r[:] for r in grid], pos[:], [OP_NULL] * n_agents
        moves_sequential(W, H, ga, pa, dead, movers, oa)
        max_rounds = max(max_rounds, moves_parallel(W, H, gb, pb, dead, movers, ob))
        assert pa == pb
        assert ga == gb
        assert oa == ob
    if bias is not None:
        assert max_rounds >= 2          # convoys: the chains really are followed
