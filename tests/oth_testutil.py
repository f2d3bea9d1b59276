"""Shared helpers for the test-suite (importable as `oth_testutil`: tests/ is on sys.path)."""
import numpy as np


def random_boards(n, seed):
    """n random (self, opp) pairs with disjoint bits at mixed densities."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 2**64, n, dtype=np.uint64)
    b = rng.integers(0, 2**64, n, dtype=np.uint64)
    c = rng.integers(0, 2**64, n, dtype=np.uint64)
    d = rng.integers(0, 2**64, n, dtype=np.uint64)
    occ = np.where(rng.random(n) < 0.5, a | b, a & b) | (c & d & a)
    s = occ & c
    o = occ & ~c
    return s, o
