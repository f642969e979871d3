"""Attractor lists are INPUT DATA of the hot path (reset, reward, step-until-attractor).  The reference obtains them
from the external CABEAN tool (utils/get_attractors_from_cabean.py:39-54, not shipped) or from its sampling recipe
`statistical_attractors` (pbn_target.py:538-560: resets x forced single updates, keep the most visited states).
This module runs that sampling recipe on the GPU and keeps the CABEAN interchange format: an attractor is a list of
cubes, a cube is a tuple over {0, 1, '*'}."""
import numpy as np
import torch

from . import engine


def parse_state(text):
    """CABEAN prints states as space-separated symbols: characters at even positions (get_attractors_from_cabean.py:9-11)."""
    return tuple(c if c == "*" else int(c) for c in text[::2] if c in "01*")


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
