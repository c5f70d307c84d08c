"""The numpy Philox4x32-10 the GPU tests predict the kernels' draws with, against the Random123 known-answer vectors
(Random123 kat_vectors, philox4x32 10 rounds)."""
import numpy as np

from philox_ref import philox4x32_10

KAT = [
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


def test_known_answer_vectors():
    for ctr, key, want in KAT:
        got = philox4x32_10([np.uint32(c) for c in ctr], [np.uint32(k) for k in key])
        assert [int(g) for g in got] == want


def test_vectorised_over_arrays():
    ctr = [np.array([c, 0], dtype=np.uint32) for c in KAT[2][0]]
    key = [np.array([k, 0], dtype=np.uint32) for k in KAT[2][1]]
    got = philox4x32_10(ctr, key)
    assert [int(g[0]) for g in got] == KAT[2][2]
    assert [int(g[1]) for g in got] == KAT[0][2]
