"""TEST INFRASTRUCTURE — imports the *unmodified* reference (read-only, /root/reference) so that
golden vectors can be recorded from it.  Runs ONLY in the build container: the GPU box has no
/root/reference, so nothing under tests/ -m gpu, smoke() or bench.py imports this module.

What it does:
  * injects throw-away stub modules for packages the image lacks (gymnasium, colomoto.types,
    matplotlib.pyplot, plotly.express) so that `import gym_PBN` succeeds;
  * applies the one-class `getState` shim: gym_PBN/envs/bittner/base.py:320-324 returns a plain
    tuple where pbn_target.py:270,314,355,385 expect an ID-keyed mapping.  The shim changes no
    arithmetic and no draw order;
  * wraps the RNG entry points the hot path uses (random.randint/random/choice/sample/uniform,
    numpy.random.uniform/rand) with a recorder, giving the replay contract of SURVEY.md §3.5.
"""
import os
import random
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("GYM_PBN_REF", "/root/reference")


# ----------------------------------------------------------------------------- stubs
class _Space:
    dtype = None


class _Discrete(_Space):
    def __init__(self, n, start=0):
        self.n, self.start = int(n), int(start)

    def contains(self, x):
        try:
            xi = int(x)
        except Exception:
            return False
        return (not isinstance(x, (list, tuple))) and self.start <= xi < self.start + self.n


class _MultiBinary(_Space):
    def __init__(self, n):
        self.n = n

    def contains(self, x):
        a = np.asarray(x)
        return a.shape == (self.n,) and bool(np.all((a == 0) | (a == 1)))


class _MultiDiscrete(_Space):
    def __init__(self, nvec):
        self.nvec = nvec

    def contains(self, x):
        return True


class _Tuple(_Space):
    def __init__(self, spaces):
        self.spaces = tuple(spaces)

    def contains(self, x):
        return (
            isinstance(x, (tuple, list))
            and len(x) == len(self.spaces)
            and all(s.contains(v) for s, v in zip(self.spaces, x))
        )


def _install_stubs():
    if "gymnasium" not in sys.modules:
        try:
            import gymnasium  # noqa: F401
        except Exception:
            gym = types.ModuleType("gymnasium")
            spaces = types.ModuleType("gymnasium.spaces")
            spaces.Discrete, spaces.MultiBinary = _Discrete, _MultiBinary
            spaces.MultiDiscrete, spaces.Tuple = _MultiDiscrete, _Tuple
            gym.spaces = spaces
            gym.Env = type("Env", (), {})
            gym._registry = {}

            def register(id, entry_point=None, **kw):
                gym._registry[id] = (entry_point, kw)

            gym.register = register
            sys.modules["gymnasium"] = gym
            sys.modules["gymnasium.spaces"] = spaces
    for name in ("colomoto", "colomoto.types", "matplotlib", "matplotlib.pyplot", "plotly", "plotly.express"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                sys.modules[name] = m
    ct = sys.modules["colomoto.types"]
    for attr in ("PartialState", "Hypercube", "TrapSpaceAttractor"):
        if not hasattr(ct, attr):
            setattr(ct, attr, type(attr, (dict,), {}))


class ShimState(tuple):
    """tuple that also answers .values()/.keys()/[gene_id] — the mapping the callers expected."""

    ids = ()

    def values(self):
        return self

    def keys(self):
        return list(self.ids)

    def __getitem__(self, k):
        if isinstance(k, (int, np.integer)) and (k >= len(self) or k < -len(self)):
            return tuple.__getitem__(self, self.ids.index(k))
        return tuple.__getitem__(self, k)


_loaded = None


def load():
    """Import the reference package (with stubs + shim) and return the module namespace we need."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError(f"reference not found at {REF_ROOT}; ref_loader is container-only")
    _install_stubs()
    sys.path.insert(0, REF_ROOT)
    # our own drop-in package has the same import name; make sure the reference wins here
    for k in [k for k in sys.modules if k == "gym_PBN" or k.startswith("gym_PBN.")]:
        del sys.modules[k]
    import gym_PBN  # noqa: F401
    from gym_PBN.envs.bittner import base

    assert os.path.realpath(base.__file__).startswith(os.path.realpath(REF_ROOT))

    def getState(self):
        st = ShimState(node.value for node in self.nodes)
        st.ids = [node.ID for node in self.nodes]
        return st

    base.Graph.getState = getState
    ns = types.SimpleNamespace()
    ns.base = base
    from gym_PBN.envs.common import pbn, pbcn, node
    from gym_PBN.envs import pbn_env, pbcn_env, pbn_target, pbn_target_multi, sampled_data
    from gym_PBN.utils import eval as ref_eval
    from gym_PBN.utils import converters, get_attractors_from_cabean

    ns.pbn, ns.pbcn, ns.node = pbn, pbcn, node
    ns.pbn_env, ns.pbcn_env, ns.pbn_target = pbn_env, pbcn_env, pbn_target
    ns.pbn_target_multi, ns.sampled_data = pbn_target_multi, sampled_data
    ns.eval, ns.converters, ns.cabean = ref_eval, converters, get_attractors_from_cabean
    _loaded = ns
    return ns


# ----------------------------------------------------------------------------- draw recorder
class Recorder:
    """Records every hot-path draw, typed: ints (randint / choice index / sample indices) and
    doubles (random / uniform / numpy uniform / numpy rand), each in call order."""

    def __init__(self):
        self.ints, self.dbls = [], []
        self._saved = None

    def take(self):
        i, d = self.ints, self.dbls
        self.ints, self.dbls = [], []
        return i, d

    def __enter__(self):
        rec = self
        inst = random._inst
        self._saved = (random.randint, random.random, random.uniform, random.choice, random.sample,
                       np.random.uniform, np.random.rand)
        o_randint, o_random, o_uniform, o_choice, o_sample, o_npu, o_nprand = self._saved

        def randint(a, b):
            v = o_randint(a, b)
            rec.ints.append(int(v))
            return v

        def rnd():
            v = o_random()
            rec.dbls.append(v)
            return v

        def uniform(a, b):
            v = o_uniform(a, b)
            rec.dbls.append(v)
            return v

        def choice(seq):
            if not len(seq):
                raise IndexError("Cannot choose from an empty sequence")
            i = inst._randbelow(len(seq))  # == CPython's Random.choice
            rec.ints.append(int(i))
            return seq[i]

        def sample(pop, k, **kw):
            res = o_sample(pop, k, **kw)
            for r in res:
                rec.ints.append(next(i for i, x in enumerate(pop) if x is r))
            return res

        def npu(low=0.0, high=1.0, size=None):
            v = o_npu(low, high, size)
            (rec.dbls.extend(np.ravel(v).tolist()) if size is not None else rec.dbls.append(float(v)))
            return v

        def nprand(*shape):
            v = o_nprand(*shape)
            rec.dbls.extend(np.ravel(v).tolist()) if shape else rec.dbls.append(float(v))
            return v

        random.randint, random.random, random.uniform = randint, rnd, uniform
        random.choice, random.sample = choice, sample
        np.random.uniform, np.random.rand = npu, nprand
        # modules that did `from random import randint` at import time
        ns = load()
        self._pbcn_randint = ns.pbcn.randint
        ns.pbcn.randint = randint
        return self

    def __exit__(self, *exc):
        (random.randint, random.random, random.uniform, random.choice, random.sample,
         np.random.uniform, np.random.rand) = self._saved
        load().pbcn.randint = self._pbcn_randint
        return False


# ----------------------------------------------------------------------------- graph builders
def build_graph(predictor_sets, node_ids):
    """Graph straight from a shipped predictor-set pickle, mirroring bittner/utils.py:81-90."""
    ns = load()
    g = ns.base.Graph(2)
    nodes = []
    for i, _id in enumerate(node_ids):
        n = ns.base.Node(i, i, f"G{_id}", _id)
        n.add_predictors(predictor_sets[i])
        nodes.append(n)
    g.add_nodes(nodes)
    return g
