import numpy as np


def booleanize(X, l):
    """Integer -> bool[l], most-significant bit first (reference: gym_PBN/utils/__init__.py:4-12)."""
    X = int(X)
    return np.array([(X >> (l - 1 - i)) & 1 for i in range(l)], dtype=bool)
