"""Synthetic networks for the stress configurations of BASELINE.json (no data files involved)."""
import numpy as np


def synthetic_pbcn(n=1024, m=8, seed=7):
    """configs[4] of BASELINE.json: M control nodes first (no inputs, P = 0), every other node three inputs
    drawn from all n nodes and the table c*f1 + (1-c)*f2 of two random Boolean functions — PBN_data 4-tuples, the summation
    order of utils/converters.py:33-34."""
    rng = np.random.default_rng(seed)
    data = []
    for i in range(n):
        mask = np.zeros(n, bool)
        if i < m:
            data.append((mask, np.array(0.0), f"u{i}", True))
            continue
        mask[rng.choice(n, size=3, replace=False)] = True
        f1, f2 = rng.integers(0, 2, 8), rng.integers(0, 2, 8)
        c = float(rng.uniform(0.1, 0.9))
        data.append((mask, (c * f1 + (1 - c) * f2).reshape(2, 2, 2), f"x{i}", False))
    return data
