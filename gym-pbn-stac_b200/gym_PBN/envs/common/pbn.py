"""Truth-table PBN core on the GPU (reference: gym_PBN/envs/common/pbn.py).

`PBN.step()` is the reference's ASYNCHRONOUS update — one node i ~ U{1..N-1} is redrawn from its truth table
(common/pbn.py:88-92; node 0 is never updated) — executed by the CUDA rollout kernel on a bit-packed device state.
`state` is exposed as a NumPy bool array like the reference's attribute (a copy: assign to `pbn.state` or call
`reset` to change it).  The exhaustive state-transition graph (`print_STG`, common/pbn.py:132-216) stays on the host:
it is O(2^N), constructor-time only, and feeds attractor lists, not the step path.
"""
import numpy as np
import torch

from gym_PBN.b200 import compiler, engine
from gym_PBN.utils import booleanize
from gym_PBN.utils.converters import logic_funcs_to_PBN_data

from .node import Node


class PBN:
    def __init__(self, PBN_data=None, logic_func_data=None, goal_config=None, device=None, seed=None):
        if PBN_data is not None and len(PBN_data) != 0:
            data = list(PBN_data)
        else:
            if logic_func_data is None:
                raise ValueError("PBN needs PBN_data or logic_func_data")
            data = logic_funcs_to_PBN_data(*logic_func_data)
        self._init_from_pbn_data(data, device, seed)
        if goal_config is not None:
            self.target_nodes = goal_config["target_nodes"]

    def _init_from_pbn_data(self, data, device=None, seed=None):
        self.spec = compiler.compile_pbn_data(data)
        self.N = self.spec.n
        self.nodes = np.empty(self.N, dtype=object)
        for i, node in enumerate(data):
            mask, table, name, ctrl = compiler._split_node(node)
            self.nodes[i] = Node(mask, table, i, self.spec.names[i], ctrl)
        self.network = engine.Network(self.spec, device=device)
        self.sim = engine.Simulator(self.network, 1, seed=0 if seed is None else seed)
        self.PBN_graph = None
        self.STG = None

    # ---- state -----------------------------------------------------------------------------------------
    @property
    def state(self):
        return self.sim.unpack()[0].cpu().numpy().astype(bool)

    @state.setter
    def state(self, value):
        bits = np.asarray(value, dtype=np.uint8).reshape(self.N) != 0
        words = np.packbits(np.pad(bits, (0, 32 * self.network.w32 - self.N)), bitorder="little").view(np.int32)
        self.sim.state.copy_(torch.from_numpy(words.copy()).view(-1, 1))  # packed on the host: one small H2D copy

    def reset(self, state=None):
        """Set the state (random when None); node 0 is forced to 0 exactly as common/pbn.py:77 does."""
        if state is None:
            new = np.random.rand(self.N) > 0.5
        else:
            if len(state) != self.N:
                raise Exception(
                    f"The length of the state given ({len(state)}) is different from the PBN size ({self.N})."
                )
            new = np.array(state, dtype=bool)
        new[0] = False
        self.state = new
        return new

    def flip(self, index):
        i = int(np.asarray(index).reshape(-1)[0])
        if not 0 <= i < self.N:
            raise IndexError(f"index {i} is out of bounds for a PBN of {self.N} nodes")
        self.sim.state[i >> 5, 0] ^= _bit32(i & 31)

    def step(self, steps=1):
        """`steps` asynchronous node updates in one kernel launch."""
        self.sim.rollout(steps)

    # ---- descriptive helpers ---------------------------------------------------------------------------
    def name_nodes(self, names):
        for node, name in zip(self.nodes, names):
            node.name = name

    def get_node_by_name(self, name):
        for node in self.nodes:
            if node.name == name:
                return node
        raise Exception(f'Node with name "{name}" not found.')

    def print_functions(self):
        return [node.function for node in self.nodes]

    def print_PBN(self, no_cache=False):
        import networkx as nx

        if self.PBN_graph is None or no_cache:
            G = nx.DiGraph()
            G.add_nodes_from(node.name for node in self.nodes)
            for node in self.nodes:
                G.add_edges_from((src.name, node.name) for src in self.nodes[node.input_mask])
            self.PBN_graph = G
        return self.PBN_graph

    def async_successors(self, state):
        """[(next_state bool[N], probability-of-node-being-1)] of the asynchronous dynamics, all N nodes considered
        (common/pbn.py:186-197: an edge exists when the node can change value)."""
        out = []
        state = np.asarray(state, dtype=bool)
        for i, node in enumerate(self.nodes):
            p = node.get_next_value_prob(state)
            if (p > 0.0 and not state[i]) or (p < 1.0 and state[i]):
                nxt = state.copy()
                nxt[i] = not state[i]
                out.append((nxt, p))
        return out

    def _compute_next_states(self, state):
        """[(str(state), str(next state), probability-of-the-node-being-1)] of the asynchronous dynamics, the edge list
        print_STG feeds to networkx (common/pbn.py:161-197, async_mode branch)."""
        state = np.asarray(state, dtype=bool)
        label = str(state.astype(int))
        return [(label, str(nxt.astype(int)), p) for nxt, p in self.async_successors(state)]

    def _probs_to_states(self, probs):
        """[(state bool[N], probability)] of the synchronous product law for a (2, N) table of per-node probabilities
        (common/pbn.py:214-232), zero-probability states left out."""
        import itertools

        out = []
        for bits in itertools.product([0, 1], repeat=probs.shape[1]):
            p = float(np.prod([probs[b, i] for i, b in enumerate(bits)]))
            if p > 0:
                out.append((np.array(bits, dtype=bool), p))
        return out

    def print_STG(self, no_cache=False):
        """networkx DiGraph over all 2^N states, node labels = str(int array) as in the reference."""
        import networkx as nx

        if self.STG is None or no_cache:
            G = nx.DiGraph()
            for idx in range(2**self.N):
                s = booleanize(idx, self.N)
                label = str(s.astype(int))
                G.add_node(label)
                G.add_weighted_edges_from((label, str(n.astype(int)), p) for n, p in self.async_successors(s))
            self.STG = G
        return self.STG

    def attractors(self, max_nodes=26):
        """Attractors of the asynchronous STG (list of sets of state tuples, as PBNEnv.compute_attractors returns them,
        pbn_env.py:238-255), computed on the device (gym_PBN.b200.attractors.exact_attractors)."""
        from gym_PBN.b200 import attractors as _att

        if self.N > max_nodes:
            raise ValueError(f"exhaustive attractor search is O(2^N); N={self.N} > {max_nodes}")
        return _att.attractor_state_sets(self.network, list_limit=1 << 22)

    def attractors_host(self, max_nodes=22):
        """Host-side reference of `attractors` (NumPy + SciPy strongly connected components), vectorised over all
        2^N states; edges follow common/pbn.py:186-197 (node i can change value: p > 0 from 0, p < 1 from 1)."""
        from scipy.sparse import csr_matrix
        from scipy.sparse.csgraph import connected_components

        if self.N > max_nodes:
            raise ValueError(f"exhaustive attractor search is O(2^N); N={self.N} > {max_nodes}")
        N, S = self.N, 2**self.N
        idx = np.arange(S, dtype=np.int64)
        bit = lambda i: (idx >> (N - 1 - i)) & 1  # state index is MSB-first like booleanize()  # noqa: E731
        rows, cols = [], []
        for i, node in enumerate(self.nodes):
            inputs = np.nonzero(node.input_mask)[0]
            tidx = np.zeros(S, dtype=np.int64)
            for j in inputs:
                tidx = (tidx << 1) | bit(j)
            p = node.function.reshape(-1)[tidx]
            cur = bit(i)
            change = ((p > 0.0) & (cur == 0)) | ((p < 1.0) & (cur == 1))
            src = idx[change]
            rows.append(src)
            cols.append(src ^ (1 << (N - 1 - i)))
        rows, cols = np.concatenate(rows), np.concatenate(cols)
        A = csr_matrix((np.ones(len(rows), np.int8), (rows, cols)), shape=(S, S))
        ncomp, label = connected_components(A, directed=True, connection="strong")
        leaves = np.ones(ncomp, bool)
        leaves[label[rows[label[rows] != label[cols]]]] = False  # a component with an edge leaving it is not terminal
        out = []
        order = np.argsort(label, kind="stable")
        bounds = np.searchsorted(label[order], np.arange(ncomp + 1))
        for comp in np.nonzero(leaves)[0]:
            members = order[bounds[comp]:bounds[comp + 1]]
            out.append({tuple(int(b) for b in booleanize(int(m), N)) for m in members})
        return out


def _bit32(b):
    """1 << b as a signed 32-bit Python int (the device planes are int32 tensors)."""
    return (1 << b) if b < 31 else -(1 << 31)
