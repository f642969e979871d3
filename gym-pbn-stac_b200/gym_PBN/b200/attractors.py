"""Attractor lists are INPUT DATA of the hot path (reset, reward, step-until-attractor).  The reference obtains them
from the external CABEAN tool (utils/get_attractors_from_cabean.py:39-54, not shipped) or from its sampling recipe
`statistical_attractors` (pbn_target.py:538-560: resets x forced single updates, keep the most visited states).
This module runs that sampling recipe on the GPU and keeps the CABEAN interchange format: an attractor is a list of
cubes, a cube is a tuple over {0, 1, '*'}."""
import numpy as np
import torch

from . import engine


def parse_state(text):
    """A state as CABEAN prints it: symbols at the even positions, '-' or '*' = don't care
    (get_attractors_from_cabean.py:9-11; the full report parser lives in gym_PBN.utils.get_attractors_from_cabean)."""
    return tuple("*" if c in "-*" else int(c) for c in text[::2])


def statistical_attractors(net, resets=100, steps=1000, top=4, care_nodes=None, seed=0):
    """Most-visited states (projected on `care_nodes` when given, other positions '*') of `resets` chains x `steps`
    forced asynchronous updates from uniformly random states.  Returns `top` single-cube attractors, most visited first."""
    sim = engine.Simulator(net, resets, seed=seed)
    sim.rand_state()
    care = np.arange(net.n) if care_nodes is None else np.asarray(care_nodes)
    seen = []
    for _ in range(steps):
        seen.append(sim.unpack()[:, care].clone())
        sim.rollout(1)
    allv = torch.cat(seen, 0)
    weights = (1 << torch.arange(min(len(care), 62), device=allv.device, dtype=torch.int64))
    if len(care) <= 62:
        keys = (allv.to(torch.int64) * weights).sum(1)
        uniq, counts = torch.unique(keys, return_counts=True)
        order = torch.argsort(counts, descending=True)[:top]
        pats = [[(int(k) >> i) & 1 for i in range(len(care))] for k in uniq[order].tolist()]
    else:
        uniq, counts = torch.unique(allv, dim=0, return_counts=True)
        order = torch.argsort(counts, descending=True)[:top]
        pats = uniq[order].cpu().numpy().astype(int).tolist()
    out = []
    for pat in pats:
        cube = ["*"] * net.n
        for pos, v in zip(care.tolist(), pat):
            cube[pos] = int(v)
        out.append([tuple(cube)])
    return out


def expand_cube(cube, limit=1 << 16):
    """All full states matching a cube (the expansion pbn_target_multi.py:438-454 builds a Python set from)."""
    stars = [i for i, v in enumerate(cube) if v == "*"]
    if 2 ** len(stars) > limit:
        raise ValueError(f"cube has {len(stars)} wildcards; refusing to expand {2**len(stars)} states")
    out = []
    for m in range(2 ** len(stars)):
        s = list(cube)
        for k, pos in enumerate(stars):
            s[pos] = (m >> (len(stars) - 1 - k)) & 1
        out.append(tuple(s))
    return out


def cube_matches(cube, state):
    return all(c == "*" or int(c) == int(s) for c, s in zip(cube, state))


# ------------------------------------------------------------------------------------------ exact attractors (N <= 32)
class StateTransitionGraph:
    """The asynchronous STG of a compiled network, held implicitly on the device: one uint32 per state with the set of
    nodes that can change value there (pbn_stg_change_masks).  Sets of states are bitsets of 2^N bits (int32 words)."""

    def __init__(self, net):
        import ctypes as C

        from . import abi

        if net.n > 32:
            raise ValueError(f"the exhaustive state-transition graph supports at most 32 nodes (got {net.n})")
        self.net, self.n = net, net.n
        self.n_states = 1 << net.n
        self.words = max(1, self.n_states >> 5)
        self._C, self._abi = C, abi
        self.masks = torch.empty(self.n_states, dtype=torch.int32, device=net.device)
        with torch.cuda.device(net.device):
            abi.check(abi.lib().pbn_stg_change_masks(net.handle, C.c_void_p(self.masks.data_ptr()), self._stream()))

    def _stream(self):
        return self._C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def empty(self):
        return torch.zeros(self.words, dtype=torch.int32, device=self.net.device)

    def full(self):
        t = torch.full((self.words,), -1, dtype=torch.int32, device=self.net.device)
        if self.n_states < 32:
            t[0] = (1 << self.n_states) - 1
        return t

    def single(self, s):
        t = self.empty()
        b = s & 31
        t[s >> 5] = (1 << b) if b < 31 else -(1 << 31)
        return t

    def pack(self, flags):
        """bool [2^N] -> bitset."""
        f = flags.to(torch.int64)
        if self.n_states < 32:
            f = torch.nn.functional.pad(f, (0, 32 - self.n_states))
        w = (f.view(-1, 32) << torch.arange(32, device=f.device, dtype=torch.int64)).sum(1)
        return torch.where(w >= (1 << 31), w - (1 << 32), w).to(torch.int32)

    @staticmethod
    def count(bits):
        x = bits.to(torch.int64) & 0xFFFFFFFF
        x = x - ((x >> 1) & 0x55555555)
        x = (x & 0x33333333) + ((x >> 2) & 0x33333333)
        x = (x + (x >> 4)) & 0x0F0F0F0F
        return int((((x * 0x01010101) & 0xFFFFFFFF) >> 24).sum())

    @staticmethod
    def first(bits):
        """Index of the lowest set bit of a non-empty bitset."""
        w = int((bits != 0).to(torch.uint8).argmax())
        v = int(bits[w]) & 0xFFFFFFFF
        return (w << 5) + ((v & -v).bit_length() - 1)

    def members(self, bits):
        """State indices (NumPy uint32) of a bitset — only for sets small enough to list."""
        nz = torch.nonzero(bits, as_tuple=False).reshape(-1)
        vals = (bits[nz].to(torch.int64) & 0xFFFFFFFF).cpu().numpy()
        out = []
        for w, v in zip(nz.cpu().numpy().tolist(), vals.tolist()):
            while v:
                low = v & -v
                out.append((w << 5) + low.bit_length() - 1)
                v ^= low
        return np.array(out, dtype=np.uint32)

    def expand(self, frontier, visited, nxt, direction, within=None):
        C = self._C
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
        with torch.cuda.device(self.net.device):
            self._abi.check(self._abi.lib().pbn_stg_expand(p(self.masks), self.n, p(frontier), p(visited), p(within), p(nxt),
                                                           direction, self._stream()))

    def reach(self, seeds, direction, within=None):
        """Bitset of the states reachable from `seeds` (direction 0) / that can reach `seeds` (direction 1), seeds included."""
        visited, frontier = seeds.clone(), seeds.clone()
        while True:
            nxt = self.empty()
            self.expand(frontier, visited, nxt, direction, within)
            if not bool(nxt.any()):
                return visited
            visited |= nxt
            frontier = nxt

    def walk(self, start, steps, seed=0):
        C = self._C
        out = torch.zeros(1, dtype=torch.int32, device=self.net.device)
        with torch.cuda.device(self.net.device):
            self._abi.check(self._abi.lib().pbn_stg_walk(C.c_void_p(self.masks.data_ptr()), self.n, int(start), int(steps),
                                                         int(seed), C.c_void_p(out.data_ptr()), self._stream()))
        return int(out.item()) & 0xFFFFFFFF


def exact_attractors(net, list_limit=1 << 16, seed=0):
    """All attractors (terminal strongly connected components of the asynchronous STG) of a network with N <= 32 nodes.

    Forward/backward reachability on device bitsets: fixed points first (every state with an empty change mask), then
    repeatedly: take a state that is still unclassified, random-walk it towards a terminal SCC, F = Fwd(p), S = Bwd(p) within F;
    F == S means F is an attractor, otherwise continue from a state of F outside S (F strictly shrinks).  After every
    attractor its whole backward closure is discarded.  Returns a list of dicts {size, states (uint32 indices, bit i = node i,
    or None above list_limit), bits (device bitset)}, largest-index-agnostic order of discovery."""
    stg = StateTransitionGraph(net)
    alive = stg.full()
    found = []

    def record(bits):
        size = stg.count(bits)
        found.append({"size": size, "states": stg.members(bits) if size <= list_limit else None, "bits": bits})

    fixed = stg.pack(stg.masks == 0)
    n_fixed = stg.count(fixed)
    if n_fixed:
        if n_fixed <= list_limit:
            for s in stg.members(fixed).tolist():
                found.append({"size": 1, "states": np.array([s], np.uint32), "bits": None})
        else:
            found.append({"size": n_fixed, "states": None, "bits": fixed, "fixed_points": True})
        alive &= ~stg.reach(fixed, 1)
    it = 0
    while bool(alive.any()):
        p = stg.walk(stg.first(alive), 64 * stg.n, seed + it)
        while True:
            it += 1
            F = stg.reach(stg.single(p), 0, within=alive)
            S = stg.reach(stg.single(p), 1, within=F)
            if torch.equal(F, S):
                break
            p = stg.walk(stg.first(F & ~S), 64 * stg.n, seed + it)
        record(F)
        alive &= ~stg.reach(F, 1, within=alive)
    return found


def attractor_state_sets(net, list_limit=1 << 16):
    """exact_attractors as a list of sets of state tuples in node order (what PBNEnv.compute_attractors returns)."""
    out = []
    for a in exact_attractors(net, list_limit):
        if a["states"] is None:
            raise ValueError(f"an attractor has {a['size']} states; too many to list (limit {list_limit})")
        out.append({tuple((int(s) >> i) & 1 for i in range(net.n)) for s in a["states"]})
    return out


def states_to_cubes(states, n):
    """Exact cover of a set of states (uint32 indices, bit i = node i) by disjoint cubes over {0, 1, '*'} in node order —
    the interchange format of the reference's attractor lists (get_attractors_from_cabean.py:14-36).  Variables under
    which the set is invariant become '*'; otherwise the set is split on the next variable (Shannon expansion)."""
    S = np.unique(np.asarray(states, dtype=np.int64))
    out = []

    def rec(S, free, fixed):
        if len(S) == 0:
            return
        stars = []
        rest = []
        cur = S
        for v in free:  # variables under which the set is invariant are wildcards
            bit = 1 << v
            if len(cur) % 2 == 0 and np.array_equal(np.sort(cur ^ bit), cur):
                stars.append(v)
                cur = cur[(cur & bit) == 0]
            else:
                rest.append(v)
        if not rest or len(cur) == 1:
            cube = ["*"] * n
            for v, val in fixed.items():
                cube[v] = val
            for v in rest:
                cube[v] = int((int(cur[0]) >> v) & 1)
            out.append(tuple(cube))
            return
        v = rest[0]
        bit = 1 << v
        lo, hi = cur[(cur & bit) == 0], cur[(cur & bit) != 0] ^ bit
        for part, val in ((lo, 0), (hi, 1)):
            f = dict(fixed)
            f[v] = val
            f.update({s: "*" for s in stars})
            rec(part, rest[1:], f)

    rec(S, list(range(n)), {})
    return out


def default_attractors(net, care_nodes, seed=0, exact_max_nodes=28):
    """Attractor list for an env that was not given one (the reference calls the external CABEAN tool here,
    pbn_target.py:530): exact terminal SCCs compressed to cubes when the network is small enough for the exhaustive
    STG and has at least two attractors, otherwise the reference's sampling recipe projected on `care_nodes`."""
    if net.n <= exact_max_nodes:
        try:
            atts = exact_attractor_cubes(net)
            if len(atts) >= 2:
                return atts, "exact"
        except ValueError:
            pass
    return statistical_attractors(net, resets=100, steps=1000, top=4, care_nodes=care_nodes, seed=seed), "sampled"


def exact_attractor_cubes(net, list_limit=1 << 20):
    """exact_attractors as the reference's list-of-cube-lists; attractors too large to list raise."""
    atts = []
    for a in exact_attractors(net, list_limit):
        if a["states"] is None:
            raise ValueError(f"an attractor has {a['size']} states; too many to compress into cubes (limit {list_limit})")
        atts.append(states_to_cubes(a["states"], net.n))
    return atts
