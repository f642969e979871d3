"""Truth-table node descriptor.  The reference's Node (gym_PBN/envs/common/node.py:5-41) evaluates itself on the
host; here it only describes one row of the compiled network — evaluation happens in the CUDA step kernel — but it
keeps `get_next_value_prob` for host-side tooling (state-transition graphs, attractor search)."""
import numpy as np


class Node:
    def __init__(self, input_mask, function, i, name=None, is_control=False):
        self.input_mask = np.asarray(input_mask, dtype=bool)
        self.function = np.asarray(function, dtype=np.float64)
        self.i = i
        self.name = name if name is not None else f"G{i}"
        self.is_control = bool(is_control)

    def value(self, state):
        return state[self.i]

    def get_next_value_prob(self, state):
        """P(next = 1 | state): table entry addressed by the masked inputs, first masked node most significant."""
        bits = np.asarray(state)[self.input_mask].astype(int)
        idx = 0
        for b in bits:
            idx = (idx << 1) | int(b)
        return float(self.function.reshape(-1)[idx])

    def compute_next_value(self, state):
        """Host-side stochastic evaluation with one draw from NumPy's legacy global generator — the reference's module does
        `from numpy import random`, so its `random.uniform(0, 1)` is numpy.random.uniform (common/node.py:2,34-38) and a seeded
        host replay consumes the same stream here.  The device step does the same comparison with Philox or replayed draws."""
        u = np.random.uniform(0, 1)
        return u < self.get_next_value_prob(state)

    def __str__(self):
        return f"{self.name}{' (Control)' if self.is_control else ''}"
