"""Army placements of the benchmark configurations (pure numpy; no engine needed).

  generate_map_positions(map_size)   the two blocks senario_battle.generate_map places (senario_battle.py:8-37)
  c4_positions()                     BASELINE config 4: 80 x 80, 512 v 512 (generate_map gives 256 v 256 there)
"""
import math

import numpy as np


def generate_map_positions(map_size):
    """-> (left [n, 3], right [n, 3]) rows of (x, y, dir): two square blocks at stride 2, a gap of 3 cells either side
    of the centre line."""
    width = height = map_size
    side = int(math.sqrt(map_size * map_size * 0.04)) * 2
    gap = 3
    ys = range((height - side) // 2, (height - side) // 2 + side, 2)
    left = [[x, y, 0] for x in range(width // 2 - gap - side, width // 2 - gap, 2) for y in ys]
    right = [[x, y, 0] for x in range(width // 2 + gap, width // 2 + gap + side, 2) for y in ys]
    return np.array(left, np.int32), np.array(right, np.int32)


def block_positions(x0, y0, cols, rows, stride=2):
    return np.array([[x0 + stride * c, y0 + stride * r, 0] for c in range(cols) for r in range(rows)], np.int32)


def c4_positions():
    """Two 16-column x 32-row blocks at stride 2, left x0 = 5, right x0 = 43, y0 = 8."""
    return block_positions(5, 8, 16, 32), block_positions(43, 8, 16, 32)
