"""Bittner predictor-graph network on the GPU (reference: gym_PBN/envs/bittner/base.py Node/Graph).

A node owns up to F predictors (3 input genes + the node's own value, linear threshold, chosen by cumulative COD
weight: base.py:30-45,89-119).  The compiler tabulates every predictor into a 16-bit LUT and the CUDA kernels do
`Graph.step` (one uniformly chosen node per call, base.py:306-312) and `Graph.synch_step` (base.py:286-303) on a
bit-packed device state.  The host objects below keep the reference's attribute surface (`graph.nodes[i].value`,
`getState`, `flipNode`, `setState`, `genRandState`, ...).
"""
import random

import numpy as np
import torch

from gym_PBN.b200 import compiler, engine


class StateView(tuple):
    """Network state as a tuple in node order that ALSO answers `.values()`, `.keys()` and `state[gene_id]`.
    The reference's getState returns a plain tuple (base.py:320-324) while its callers index it by gene ID
    (pbn_target.py:270,355,385); this type serves both uses."""

    ids = ()

    def values(self):
        return self

    def keys(self):
        return list(self.ids)

    def items(self):
        return list(zip(self.ids, self))

    def __getitem__(self, k):
        if isinstance(k, (int, np.integer)) and not (-len(self) <= k < len(self)):
            return tuple.__getitem__(self, self.ids.index(int(k)))
        return tuple.__getitem__(self, k)


class Node:
    """Descriptor of one gene.  `.value` reads / writes the live device state through the owning graph."""

    def __init__(self, index, bittnerIndex, name, ID, LUTflag=False):
        self.index, self.bittnerIndex, self.name, self.ID, self.LUTflag = index, bittnerIndex, name, ID, LUTflag
        self.CODsum = 0
        self.predictors = []
        self._raw = None
        self._graph = None
        self._value = None

    def add_predictors(self, predictors):
        """(3, F) object array of (COD, A(4,1), input IDs) columns; cumulative COD as base.py:30-45."""
        self._raw = predictors
        for COD, A, inputIDs in np.asarray(predictors, dtype=object).T:
            if COD is None:
                continue
            self.CODsum += COD
            prev = self.predictors[-1][2] if self.predictors else 0
            self.predictors.append((inputIDs, A, COD if not self.predictors else prev + COD))

    def getInputNodes(self):
        seen = []
        for ids, _A, _c in self.predictors:
            for g in ids:
                if g not in seen:
                    seen.append(g)
        return seen

    @property
    def value(self):
        if self._graph is None:
            return self._value
        return int(self._graph._bits()[self.index])

    @value.setter
    def value(self, v):
        if self._graph is None:
            self._value = v
        else:
            self._graph._set_bit(self.index, int(v))

    def setValue(self, value):
        self.value = value


class Graph:
    def __init__(self, base=2, device=None, seed=None):
        self.nodes = []
        self.edges = []
        self.base = base
        self.perturbations = False
        self.p = 0.001
        self.is_directed = True
        self._device, self._seed = device, seed
        self.network = None
        self.sim = None
        self.spec = None

    # ---- construction ---------------------------------------------------------------------------------
    @classmethod
    def from_spec(cls, spec, device=None, seed=None):
        g = cls(2, device, seed)
        g.spec = spec
        g.nodes = [Node(i, i, spec.names[i], spec.ids[i]) for i in range(spec.n)]
        g._bind()
        return g

    def add_nodes(self, nodeList):
        """Nodes built the reference way (Node(...).add_predictors(pickle[i]), bittner/utils.py:81-90)."""
        self.nodes = list(nodeList)
        sets = [n._raw for n in self.nodes]
        self.spec = compiler.compile_predictor_sets(sets, [n.ID for n in self.nodes], [str(n.name) for n in self.nodes])
        self._bind()

    def _bind(self):
        self.network = engine.Network(self.spec, device=self._device)
        self.sim = engine.Simulator(self.network, 1, seed=0 if self._seed is None else self._seed)
        for n in self.nodes:
            n._graph = self
        self._ids = [n.ID for n in self.nodes]

    @property
    def N(self):
        return len(self.nodes)

    # ---- state ------------------------------------------------------------------------------------------
    def _bits(self):
        return self.sim.unpack()[0].cpu().numpy()

    def _set_bit(self, i, v):
        bits = self._bits()
        bits[i] = v
        self.sim.set_state(bits.reshape(1, -1))

    def state_view(self, bits):
        st = StateView(int(b) for b in bits)
        st.ids = self._ids
        return st

    def getState(self):
        return self.state_view(self._bits())

    def getLabeledState(self):
        return dict(zip((n.ID for n in self.nodes), (int(b) for b in self._bits())))

    def setState(self, state):
        bits = np.array([int(v) for v in state], dtype=np.uint8)
        words = np.packbits(np.pad(bits, (0, 32 * self.network.w32 - self.N)), bitorder="little").view(np.int32)
        self.sim.state.copy_(torch.from_numpy(words.copy()).view(-1, 1))  # packed on the host: one small H2D copy

    def genRandState(self):
        """randint(0, base-1) per node from Python's `random`, as base.py:368-370."""
        self.setState([random.randint(0, self.base - 1) for _ in range(self.N)])

    def flipNode(self, index):
        if index < len(self.nodes) and index >= 0:
            self.sim.state[index >> 5, 0] ^= (1 << (index & 31)) if (index & 31) < 31 else -(1 << 31)
        else:
            raise ValueError(f"Invalid action, no node at index {index}")

    # ---- dynamics ---------------------------------------------------------------------------------------
    def step(self, changed_nodes=None, i=None, steps=1):
        """Asynchronous update(s) on the device; returns the new state like base.py:306-312."""
        if i is not None:
            raise NotImplementedError("forcing the updated node index is not supported on the device path")
        self.sim.rollout(steps)
        return self.getState()

    def synch_step(self):
        if self.perturbations:
            raise NotImplementedError("synch_step with perturbations (base.py:287-299) is not on the device path")
        self.sim.rollout(1, sync=True)

    # ---- descriptive ------------------------------------------------------------------------------------
    def getNames(self):
        return [[n.name] for n in self.nodes]

    def getIDs(self):
        return [n.ID for n in self.nodes]

    def getNodeByID(self, ID):
        for n in self.nodes:
            if n.ID == ID:
                return n
        return None
