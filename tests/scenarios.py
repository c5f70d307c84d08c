"""Deterministic placements and action streams shared by tests, golden generation and bench
(SURVEY.md section 8d).  Pure numpy; no engine needed."""
import numpy as np


from mfmarl_b200.scenarios import block_positions, c4_positions, generate_map_positions  # noqa: F401  (the product's own placements)


# action ids for the battle config: 0-12 moves, 13-20 attacks
MOVE_DELTAS = [(0, -2), (-1, -1), (0, -1), (1, -1), (-2, 0), (-1, 0), (0, 0), (1, 0), (2, 0),
               (-1, 1), (0, 1), (1, 1), (0, 2)]


def fight_actions(rng, pos, map_size):
    """50 % random attack / 35 % advance toward the map centre / 15 % uniform -- produces kills."""
    n = len(pos)
    u = rng.random_sample(n)
    acts = rng.randint(0, 21, size=n)
    attack = rng.randint(13, 21, size=n)
    c = map_size // 2
    dx = np.sign(c - pos[:, 0]).astype(int)
    dy = np.sign(c - pos[:, 1]).astype(int)
    toward = np.array([MOVE_DELTAS.index((int(a), int(b))) for a, b in zip(dx, dy)], dtype=int) \
        if n else np.zeros((0,), int)
    out = np.where(u < 0.5, attack, np.where(u < 0.85, toward, acts))
    return out.astype(np.int32)


def uniform_actions(rng, n):
    return rng.randint(0, 21, size=n).astype(np.int32)


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
