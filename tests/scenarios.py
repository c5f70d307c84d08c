"""Deterministic placements and action streams shared by tests, golden generation and bench
(SURVEY.md section 8d).  Pure numpy; no engine needed."""
import numpy as np


def generate_map_positions(map_size):
    """Placement of senario_battle.generate_map (senario_battle.py:8-37): two square blocks at
    stride 2, `gap` 3 either side of the centre line.  -> (left [n,3], right [n,3])."""
    import math
    width = height = map_size
    side = int(math.sqrt(map_size * map_size * 0.04)) * 2
    gap = 3
    left = [[x, y, 0] for x in range(width // 2 - gap - side, width // 2 - gap, 2)
            for y in range((height - side) // 2, (height - side) // 2 + side, 2)]
    right = [[x, y, 0] for x in range(width // 2 + gap, width // 2 + gap + side, 2)
             for y in range((height - side) // 2, (height - side) // 2 + side, 2)]
    return np.array(left, np.int32), np.array(right, np.int32)


def block_positions(x0, y0, cols, rows, stride=2):
    return np.array([[x0 + stride * c, y0 + stride * r, 0] for c in range(cols) for r in range(rows)],
                    np.int32)


def c4_positions():
    """BASELINE config 4 (80x80, 512 v 512): two 16-col x 32-row blocks at stride 2, left x0=5,
    right x0=43, y0=8."""
    return block_positions(5, 8, 16, 32), block_positions(43, 8, 16, 32)


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
