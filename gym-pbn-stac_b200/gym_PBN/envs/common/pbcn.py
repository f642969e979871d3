"""PBCN core = PBN + a stored control vector (reference: gym_PBN/envs/common/pbcn.py).

In this fork `apply_control` only stores `control_state`; the asynchronous `step` reads `self.state` alone, so the
control never reaches the dynamics (common/pbcn.py:40-49,59-61 — quirk Q4 of SURVEY.md).  `control="stac"` (default)
reproduces that; `control="write"` writes the control vector into state[0:M] before each update, the semantics the
upstream code implies (common/pbcn.py:54-58).
"""
import numpy as np

from gym_PBN.b200 import compiler

from .node import Node
from .pbn import PBN


class PBCN(PBN):
    def __init__(self, PBN_data=None, logic_func_data=None, device=None, seed=None, control="stac"):
        super().__init__(PBN_data, logic_func_data, device=device, seed=seed)
        if control not in ("stac", "write"):
            raise ValueError("control must be 'stac' (reference behaviour) or 'write'")
        self.control_mode = control
        flags = self.spec.control
        self.control_nodes = np.array([n for n, f in zip(self.nodes, flags) if f], dtype=object)
        self.M = len(self.control_nodes)
        self.control_state = np.zeros(self.M, dtype=bool)

    def apply_control(self, control):
        if len(control) != len(self.control_nodes):
            raise ValueError(
                f"Control for {len(control)} control nodes provided, when there are {len(self.control_nodes)} in the network."
            )
        self.control_state = np.array(control, dtype=bool)

    def step(self, steps=1):
        if self.control_mode == "write" and self.M:
            for _ in range(steps):
                s = self.state
                s[: self.M] = self.control_state
                self.state = s
                self.sim.rollout(1)
        else:
            self.sim.rollout(steps)

    @property
    def control_actions(self):
        """Every 0/1 vector of length N (common/pbcn.py:114-116 — over N, not M, as there)."""
        import itertools

        return map(list, itertools.product([0, 1], repeat=self.N))

    def _async_compute_next_states(self, state):
        """Edge list of the asynchronous STG with the control vector prepended to the target labels' source array, as
        common/pbcn.py:72-93 builds it (the node probabilities read the plain state)."""
        state = np.asarray(state, dtype=bool)
        combined = np.concatenate((self.control_state, state))
        out = []
        for i in range(self.N):
            p = self.nodes[i].get_next_value_prob(state)
            if (p > 0.0 and not state[i]) or (p < 1.0 and state[i]):
                nxt = combined.copy()
                nxt[i] = not state[i]
                out.append((str(state.astype(int)), str(nxt.astype(int)), p))
        return out

    def reset(self, state=None):
        self.control_state = np.zeros(self.M, dtype=bool)
        return super().reset(state=state)
