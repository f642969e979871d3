"""Helpers shared by the oracle tests (CPU) and the CUDA parity tests (GPU): load a golden trace
recorded from the reference (oracle/make_golden.py) and build oracle-side networks/envs from it."""
from pathlib import Path

import numpy as np

GOLD = Path(__file__).resolve().parent / "golden"


def load(name):
    return np.load(GOLD / name, allow_pickle=False)


class Traj:
    def __init__(self, z, e):
        p = f"e{e}/"
        for k in ("op", "act", "int_off", "ints", "dbl_off", "dbls", "state", "obs", "reward", "term", "trunc", "target_att"):
            setattr(self, k, z[p + k])
        self.T = len(self.op)

    def draws(self, t, t1=None):
        t1 = t + 1 if t1 is None else t1
        return (self.ints[self.int_off[t]:self.int_off[t1]], self.dbls[self.dbl_off[t]:self.dbl_off[t1]])


def pbn_data_from(z):
    """(input_mask, table[2]*k) tuples back from the padded arrays stored in a truth-table golden."""
    out = []
    for m, t in zip(z["masks"], z["tables"]):
        k = int(m.sum())
        out.append((m.astype(bool), t[: 2**k].reshape([2] * k) if k else np.array(t[0])))
    return out


def cubes_to_attractors(cubes, off):
    atts = []
    for a in range(len(off) - 1):
        atts.append([tuple("*" if v == 2 else int(v) for v in c) for c in cubes[off[a]:off[a + 1]]])
    return atts


from gym_PBN.b200.synthetic import synthetic_pbcn  # noqa: E402,F401  (configs[4]: lives in the package, bench.py uses it too)
