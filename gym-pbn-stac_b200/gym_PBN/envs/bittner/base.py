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

    # ---- host-side analysis helpers (the reference's STG / attractor side paths; not on the device hot path) -----------
    def addInputNode(self, inputNode):
        self.inputNodes = list(getattr(self, "inputNodes", [])) + [inputNode]

    def addInputNodes(self, inputNodes):
        self.inputNodes = list(getattr(self, "inputNodes", [])) + list(inputNodes)

    def addLUT(self, LUT, inputIDs):
        self.LUT, self.inputIDs = LUT, inputIDs

    def _table(self):
        """[(input node indices (3 + self), 16-bit LUT, cumulative COD)] and CODsum of this node, from the compiled network."""
        if self._graph is None:
            raise Exception("node is not part of a graph yet (Graph.add_nodes)")
        a = self._graph.spec.arrays
        q0, q1 = int(a["pr_off"][self.index]), int(a["pr_off"][self.index + 1])
        rows = [(a["pr_in"][4 * q:4 * q + 4].tolist(), int(a["pr_lut"][q]), float(a["pr_cum"][q])) for q in range(q0, q1)]
        return rows, float(a["pr_codsum"][self.index])

    def _output(self, state, inputs, lut):
        g = self._graph
        bits = [g._state_bit(state, j) for j in inputs]
        return (lut >> ((bits[0] << 3) | (bits[1] << 2) | (bits[2] << 1) | bits[3])) & 1

    def getStateProbs(self, state):
        """[P(next = 0), P(next = 1)] given a state (base.py:66-87): each predictor votes with its COD share."""
        rows, codsum = self._table()
        probs, prev = [0, 0], 0.0
        for inputs, lut, cum in rows:
            probs[self._output(state, inputs, lut)] += (cum - prev) / codsum
            prev = cum
        return probs

    def Predstep(self, state):
        """One host-side evaluation with a draw from Python's `random`, as base.py:89-119 (the device path does the same
        with Philox or replayed draws)."""
        rows, codsum = self._table()
        r = random.random() * codsum
        chosen = rows[-1]
        for row in rows:
            if row[2] > r:
                chosen = row
                break
        return int(self._output(state, chosen[0], chosen[1]))

    def LUTstep(self, state):
        raise ValueError("you shouldn't be here")

    def step(self, state, verbose=False):
        Y = self.LUTstep(state) if self.LUTflag else self.Predstep(state)
        self.value = Y
        return Y


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
        """Asynchronous update(s) on the device; returns the new state like base.py:306-312.  With a forced node index `i`
        the single update is evaluated on the host (Node.step, one draw from Python's `random`)."""
        if i is not None:
            self.nodes[i].step(self.getLabeledState())
            return self.getState()
        self.sim.rollout(steps)
        return self.getState()

    def synch_step(self):
        if self.perturbations:
            raise NotImplementedError("synch_step with perturbations (base.py:287-299) is not on the device path")
        self.sim.rollout(1, sync=True)

    # ---- state-transition analysis (host side; reference: base.py:199-254,398-399) -------------------------
    def _state_bit(self, state, j):
        """Value of node index j in `state`: a sequence in node order, or a mapping keyed by gene ID."""
        if isinstance(state, dict):
            return int(state[self.nodes[j].ID])
        return int(state[j])

    def _moves(self, votes_at, start):
        from collections import defaultdict

        probs = [node.getStateProbs(votes_at) for node in self.nodes]
        nextStates = defaultdict(float)
        for i in range(len(start)):
            nextState = list(start)
            for v in (0, 1):
                if probs[i][v] > 0.0:
                    nextState[i] = v
                    nextStates[tuple(nextState)] += probs[i][v] / len(start)
        return nextStates

    def getNextStates(self, state=None):
        """{next state tuple: probability} of the asynchronous dynamics: node i (chosen w.p. 1/N) takes each value its
        predictors give with positive probability.  The predictor votes are taken at the graph's CURRENT state, the move
        starts from `state` (current state when omitted) — as base.py:221-242 does."""
        cur = tuple(self.getState())
        return self._moves(cur, cur if state is None else tuple(state))

    def sync_getNextStates(self):
        """{next state tuple: probability} of the synchronous dynamics (base.py:244-259): product over nodes."""
        import itertools
        from collections import defaultdict

        cur = self.getState()
        probs = [node.getStateProbs(cur) for node in self.nodes]
        nextStates = defaultdict(float)
        for state in itertools.product([0, 1], repeat=len(self.nodes)):
            p = 1
            for i, v in enumerate(state):
                p *= probs[i][v]
            if p > 0:
                nextStates[state] = p
        return nextStates

    def genSTG(self, savepath=None):
        """networkx DiGraph of the asynchronous state-transition graph over all 2^N state tuples (base.py:199-218),
        including the self-loops the reference adds for values a node can keep.  Host side, for small networks; the
        attractors of networks up to 32 nodes come from the device search (getAttractors / b200.attractors)."""
        import itertools
        import os
        import pickle

        import networkx as nx

        if savepath is not None and os.path.exists(savepath):
            with open(savepath, "rb") as f:
                return pickle.load(f)
        if self.N > 16:
            raise ValueError(f"the explicit state-transition graph has 2^{self.N} nodes; use getAttractors (N <= 32) instead")
        stg = nx.DiGraph()
        states = list(itertools.product([0, 1], repeat=self.N))
        stg.add_nodes_from(states)
        for s in states:
            for nxt in self._moves(s, s):
                stg.add_edge(s, nxt)
        if savepath is not None:
            with open(savepath, "wb") as f:
                pickle.dump(stg, f)
        return stg

    def getAttractors(self, verbal=False):
        """Attractors (terminal strongly connected components of the asynchronous STG) as lists of state tuples, from the
        exhaustive device search.  (The reference's method calls an undefined `attractorSetFinder`, base.py:372-374.)"""
        from gym_PBN.b200 import attractors as _att

        return [sorted(a) for a in _att.attractor_state_sets(self.network)]

    def addEdge(self, startIndex, endIndex):
        self.nodes[endIndex].addInputNode(self.nodes[startIndex])
        self.edges = self.edges + [(self.nodes[startIndex], self.nodes[endIndex])]

    def addCon(self, conn):
        k = conn.shape[0] - 1
        self.k = k
        for i in range(conn.shape[1]):
            targNode = self.getNodeByID(conn[k, i])
            for j in range(k):
                predNode = self.getNodeByID(conn[j, i])
                self.edges = self.edges + [(predNode.index, targNode.index)]
                targNode.addInputNode(predNode)

    def printGraph(self, path=None, dist=10, charLim=10):
        """networkx DiGraph of the gene connectivity, node labels = gene names cut to charLim (base.py:341-362)."""
        import networkx as nx

        self.G = nx.DiGraph()
        label = lambda node: str(node.name)[:charLim]  # noqa: E731
        for node in self.nodes:
            self.G.add_node(label(node))
            for ID in node.getInputNodes():
                self.G.add_edge(label(self.getNodeByID(ID)), label(node))
        return self.G

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


def genBoolList(n, length, b):
    output = np.zeros((length), dtype=int)
    for i in range(length):
        output[i] = n % b
        n = n // b
    return output


def integerize(state):
    return sum(int(v) * (2**i) for i, v in enumerate(state))


def findAttractors(stg):
    """Attracting components of a state-transition graph (base.py:398-399)."""
    import networkx as nx

    return list(nx.attracting_components(stg))
