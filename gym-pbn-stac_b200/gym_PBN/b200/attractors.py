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
    STG and has at least two attractors; beyond that the sampled + verified route (closed cube sets); the reference's
    sampling recipe projected on `care_nodes` (no closure guarantee) only when that finds fewer than two attractors."""
    if net.n <= exact_max_nodes:
        try:
            atts = exact_attractor_cubes(net)
            if len(atts) >= 2:
                return atts, "exact"
        except ValueError:
            pass
    try:  # sampling + verification: closed cube sets (trap spaces, cut down to exact SCCs where they can be enumerated)
        atts, _ = verified_attractors(net, seed=seed)
        if len(atts) >= 2:
            return atts, "verified"
    except (ValueError, RuntimeError):
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


# ------------------------------------------------------------------------ sampled + VERIFIED attractors (any N)
# Beyond ~32 nodes the state-transition graph cannot be enumerated, and the reference hands the job to the external CABEAN
# tool (utils/get_attractors_from_cabean.py:39-54).  Here: sample states by long asynchronous rollouts on the GPU, grow each
# sampled state into the smallest TRAP SPACE (a cube closed under every possible update) that contains it, and — when that
# cube is small enough to enumerate — cut it down to the exact terminal strongly connected components inside it.  Every
# cube set that is returned has been checked symbolically: each successor of each of its cubes stays inside the set.
class SuccessorModel:
    """What an asynchronous update of node i can produce on a cube: host-side view of a compiled NetworkSpec."""

    def __init__(self, spec):
        from . import abi

        self.spec, self.n, self.first = spec, spec.n, spec.first_updatable
        a = spec.arrays
        self.pred = spec.kind == abi.NET_PRED
        if self.pred:
            off, cum = a["pr_off"], a["pr_cum"]
            self.rows = []  # per node: list of (inputs[4], lut16) of the predictors that can be selected
            for i in range(self.n):
                row, prev = [], 0.0
                for k in range(off[i], off[i + 1]):
                    if cum[k] > prev:  # positive selection weight (bittner/base.py:94-97: first cum_k > r)
                        row.append((tuple(int(v) for v in a["pr_in"][4 * k:4 * k + 4]), int(a["pr_lut"][k])))
                    prev = max(prev, float(cum[k]))
                self.rows.append(row)
        else:
            self.rows = []
            for i in range(self.n):
                ins = tuple(int(v) for v in a["tt_in"][a["tt_in_off"][i]:a["tt_in_off"][i + 1]])
                prob = np.asarray(a["tt_prob"][a["tt_tab_off"][i]:a["tt_tab_off"][i + 1]], np.float64)
                self.rows.append([(ins, prob)])

    def outcomes(self, i, cube):
        """-> list of (new value b, {input node: value} that the source state must satisfy) over every way node i can be
        updated from a state of `cube` (a sequence over {0, 1, '*'}); identical constraints are merged."""
        out = {}
        for ins, table in self.rows[i]:
            k = len(ins)
            for idx in range(1 << k):
                req, ok = {}, True
                for q, node in enumerate(ins):  # first input = most significant bit of the table index
                    bit = (idx >> (k - 1 - q)) & 1
                    if cube[node] != "*" and int(cube[node]) != bit:
                        ok = False
                        break
                    if req.get(node, bit) != bit:  # the same node twice among the inputs with different bits
                        ok = False
                        break
                    req[node] = bit
                if not ok:
                    continue
                if self.pred:
                    vals = ((table >> idx) & 1,)
                else:
                    p = float(table[idx])
                    vals = tuple(b for b, possible in ((0, p < 1.0), (1, p > 0.0)) if possible)
                for b in vals:
                    out.setdefault((b, tuple(sorted(req.items()))), None)
        return [(b, dict(req)) for (b, req) in out]

    def can(self, i, cube):
        """(node i can become 0, node i can become 1) from some state of the cube."""
        vals = {b for b, _ in self.outcomes(i, cube)}
        return 0 in vals, 1 in vals


def trap_space(model, state):
    """Smallest cube that contains `state` and is closed under every asynchronous update: free every fixed node that some
    update can flip, until nothing changes (percolation).  Nodes below first_updatable never change (common/pbn.py:90)."""
    cube = [int(v) for v in state]
    changed = True
    while changed:
        changed = False
        for i in range(model.first, model.n):
            if cube[i] == "*":
                continue
            c0, c1 = model.can(i, cube)
            if (c1 if cube[i] == 0 else c0):
                cube[i] = "*"
                changed = True
    return tuple(cube)


def _cube_in_union(d, cubes):
    """Is cube d contained in the union of `cubes`?  (split d on a variable some candidate fixes)"""
    cands = [c for c in cubes if all(cv == "*" or dv == "*" or cv == dv for cv, dv in zip(c, d))]  # those that meet d
    for c in cands:
        if all(cv == "*" or cv == dv for cv, dv in zip(c, d)):
            return True
    for c in cands:
        for v, (cv, dv) in enumerate(zip(c, d)):
            if dv == "*" and cv != "*":
                lo, hi = list(d), list(d)
                lo[v], hi[v] = 0, 1
                return _cube_in_union(tuple(lo), cands) and _cube_in_union(tuple(hi), cands)
    return False


def cubes_closed(model, cubes):
    """Verification: every successor of every state of every cube lies in the union of the cubes."""
    cubes = [tuple(c) for c in cubes]
    for c in cubes:
        for i in range(model.first, model.n):
            for b, req in model.outcomes(i, c):
                if c[i] == "*" or c[i] == b:
                    continue  # the successor stays in c
                d = list(c)
                for node, bit in req.items():
                    d[node] = bit
                if d[i] != "*" and d[i] != c[i]:
                    continue  # the constraint contradicts the cube on node i itself (its own value is an input)
                d[i] = b
                if not _cube_in_union(tuple(d), cubes):
                    return False
    return True


def terminal_sccs_in_cube(model, cube, max_free=18):
    """Exact attractors inside a closed cube with at most max_free wildcards: explicit asynchronous STG over the 2^f states,
    strongly connected components (SciPy), those without an outgoing edge.  Returns a list of cube lists."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components

    free = [v for v, c in enumerate(cube) if c == "*"]
    f = len(free)
    if f > max_free:
        raise ValueError(f"{f} wildcards: too many states to enumerate")
    pos = {v: k for k, v in enumerate(free)}
    idx = np.arange(1 << f, dtype=np.int64)

    def bit_of(node):  # value of `node` in every enumerated state
        return ((idx >> pos[node]) & 1) if node in pos else np.full(idx.shape, int(cube[node]), np.int64)

    src, dst = [], []
    for i in free:
        if i < model.first:
            continue
        cur = bit_of(i)
        for ins, table in model.rows[i]:
            k = len(ins)
            t = np.zeros(idx.shape, np.int64)
            for node in ins:
                t = (t << 1) | bit_of(node)
            if model.pred:
                outs = [((table >> t) & 1)]
            else:
                p = np.asarray(table)[t]
                outs = [np.where(p > 0.0, 1, cur), np.where(p < 1.0, 0, cur)]
            for o in outs:
                ch = o != cur
                src.append(idx[ch])
                dst.append(idx[ch] ^ (1 << pos[i]))
    S = np.concatenate(src) if src else np.zeros(0, np.int64)
    D = np.concatenate(dst) if dst else np.zeros(0, np.int64)
    g = coo_matrix((np.ones(len(S), np.int8), (S, D)), shape=(1 << f, 1 << f)).tocsr()
    ncomp, lab = connected_components(g, directed=True, connection="strong")
    leaves = np.ones(ncomp, bool)
    leaves[lab[S[lab[S] != lab[D]]]] = False
    out = []
    for comp in np.nonzero(leaves)[0]:
        members = idx[lab == comp]
        cubes = []
        for sub in states_to_cubes(members, f):  # over the free variables; embed into the full cube
            full = list(cube)
            for k, v in enumerate(free):
                full[v] = sub[k]
            cubes.append(tuple(full))
        out.append(cubes)
    return out


def restricted_network(spec, cube):
    """The network a closed cube induces on its wildcard nodes (predictor networks): fixed inputs are folded into the LUTs,
    so the exhaustive device search (exact_attractors) can run inside a trap space of a network of any size.
    -> (NetworkSpec over the free nodes, list of their original indices)."""
    from . import abi
    from .compiler import NetworkSpec

    if spec.kind != abi.NET_PRED:
        raise ValueError("restricted_network handles predictor networks")
    a = spec.arrays
    free = [v for v, c in enumerate(cube) if c == "*"]
    pos = {v: k for k, v in enumerate(free)}
    off, ins, luts, cums, sums = [0], [], [], [], []
    for i in free:
        for k in range(a["pr_off"][i], a["pr_off"][i + 1]):
            src = [int(v) for v in a["pr_in"][4 * k:4 * k + 4]]
            lut, new = int(a["pr_lut"][k]), 0
            for idx in range(16):
                full = 0
                for q, node in enumerate(src):
                    bit = (idx >> (3 - q)) & 1 if node in pos else int(cube[node])
                    full |= bit << (3 - q)
                new |= ((lut >> full) & 1) << idx
            ins += [pos.get(node, 0) for node in src]  # a folded input points anywhere: the LUT no longer depends on it
            luts.append(new)
            cums.append(float(a["pr_cum"][k]))
        sums.append(float(a["pr_codsum"][i]))
        off.append(len(luts))
    arrays = dict(pr_off=np.array(off, np.int32), pr_in=np.array(ins, np.int32), pr_lut=np.array(luts, np.uint16),
                  pr_cum=np.array(cums, np.float64), pr_codsum=np.array(sums, np.float64))
    return NetworkSpec(abi.NET_PRED, len(free), 0, [spec.names[v] for v in free], arrays=arrays), free


def terminal_sccs_on_device(net, cube, list_limit=1 << 22):
    """Exact attractors inside a closed cube with at most 28 wildcards, by the exhaustive device search on the restricted
    network.  Returns a list of cube lists (full-length cubes)."""
    sub, free = restricted_network(net.spec, cube)
    out = []
    for att in exact_attractors(engine.Network(sub, device=net.device), list_limit):
        if att["states"] is None:
            raise ValueError("an attractor inside the trap space is too large to list")
        cubes = []
        for c in states_to_cubes(att["states"], len(free)):
            full = list(cube)
            for k, v in enumerate(free):
                full[v] = c[k]
            cubes.append(tuple(full))
        out.append(cubes)
    return out


def verified_attractors(net, resets=256, steps=None, seed=0, max_free_host=14, max_free_device=26):
    """Attractors by sampling + verification, for networks of any size: `resets` uniformly random states run `steps`
    asynchronous updates on the GPU; each distinct end state is grown into its smallest trap space (smallest first: a walker
    that has not converged yet gives a larger one that CONTAINS attractors already found, and is dropped when it cannot be
    enumerated); trap spaces small enough to enumerate are replaced by the exact terminal SCCs inside them — on the host up to
    max_free_host wildcards, by the exhaustive device search on the restricted network up to max_free_device.  Every returned
    cube set passed `cubes_closed`.
    -> (attractors: list of cube lists, info: list of {"method", "free", "states"} per attractor)."""
    from . import abi

    steps = int(steps) if steps else 512 * net.n  # slow transients: a 200-node set needs ~1e5 updates to settle
    sim = engine.Simulator(net, int(resets), seed=seed)
    sim.rand_state()
    sim.rollout(steps)
    ends = np.unique(sim.unpack().cpu().numpy(), axis=0)
    model = SuccessorModel(net.spec)
    spaces = set()
    for s in ends:
        if not any(all(c == "*" or c == int(v) for c, v in zip(t, s)) for t in spaces):
            spaces.add(trap_space(model, s))  # (a state inside a known trap space has its own inside it too: found by refinement)
    atts, info, seen = [], [], set()

    def contains(big, small):
        return all(b == "*" or b == c for b, c in zip(big, small))

    for t in sorted(spaces, key=lambda c: sum(v == "*" for v in c)):
        free = sum(v == "*" for v in t)
        if free <= max_free_host:
            found = [(c, "exact terminal SCC inside a sampled trap space (host)") for c in terminal_sccs_in_cube(model, t, max_free_host)]
        elif free <= max_free_device and net.spec.kind == abi.NET_PRED:
            found = [(c, "exact terminal SCC inside a sampled trap space (device search on the restricted network)")
                     for c in terminal_sccs_on_device(net, t)]
        elif any(contains(t, c) for cubes in atts for c in cubes):
            continue  # an unconverged sample: the cube holds an attractor we already have and cannot be searched for more
        else:
            found = [([t], "sampled trap space (closed cube; too large to enumerate)")]
        for cubes, method in found:
            key = tuple(sorted(cubes, key=str))
            if key in seen:
                continue
            if not cubes_closed(model, cubes):
                raise RuntimeError("internal error: an attractor candidate is not closed under the dynamics")
            seen.add(key)
            atts.append(list(cubes))
            info.append({"method": method, "free": free, "states": sum(2 ** sum(v == "*" for v in c) for c in cubes)})
    return atts, info
