"""Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11) in numpy, written
from the published round function; test infrastructure for the kernels' counter-based draws (csrc/rng.cuh)."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(ctr, key, rounds=10):
    """ctr: four uint32 arrays, key: two uint32 arrays (broadcastable) -> four uint32 arrays"""
    c = [np.asarray(a, dtype=np.uint32) for a in ctr]
    k = [np.asarray(a, dtype=np.uint32) for a in key]
    lo32 = np.uint64(0xFFFFFFFF)
    for _ in range(rounds):
        p0 = M0 * c[0].astype(np.uint64)
        p1 = M1 * c[2].astype(np.uint64)
        hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & lo32).astype(np.uint32)
        hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & lo32).astype(np.uint32)
        c = [hi1 ^ c[1] ^ k[0], lo1, hi0 ^ c[3] ^ k[1], lo0]
        with np.errstate(over="ignore"):
            k = [(k[0] + W0).astype(np.uint32), (k[1] + W1).astype(np.uint32)]
    return c
