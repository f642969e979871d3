"""Network compiler, front end (host, NumPy): reference-form network descriptions -> flat arrays of
`PbnNetDesc` (include/pbn_b200.h).  The back end (pbn_net_create, C++) lowers these into the packed
device image: u8/u16 gather indices, 16-bit LUTs and 31-bit integer thresholds.

Front ends
  * truth tables  — PBN_data 4-tuples (mask, table, name, control) as produced by
    logic_funcs_to_PBN_data, or README-style 5-tuples (mask, table, i, name, is_control)
    (reference: common/pbn.py:37-46, common/node.py:6-32, common/pbcn.py:22-28; quirk Q5).
  * predictor sets — the shipped pickles: per node a (3, F) object array of (COD, A(4,1), input IDs)
    (reference: bittner/base.py:30-45 add_predictors, :89-119 Predstep).  The 16-entry LUT of each
    predictor is tabulated HERE with the reference's own float expression `np.matmul(X.T, A) < 0.`,
    because a handful of entries sit exactly at / within rounding of zero; the GPU
    never redoes that dot product.
  * attractor / target cubes — tuples over {0, 1, '*'} (the CABEAN interchange format,
    utils/get_attractors_from_cabean.py:14-36) -> int8 arrays with 2 for '*'.
"""
import json
import pickle
from dataclasses import dataclass, field
from pathlib import Path
from typing import List, Optional

import numpy as np

from . import abi

DATA_DIR = Path(__file__).resolve().parents[1] / "envs" / "bittner" / "data"


@dataclass
class NetworkSpec:
    kind: int
    n: int
    first_updatable: int
    names: List[str]
    ids: Optional[List[int]] = None          # gene IDs (predictor graphs)
    control: Optional[np.ndarray] = None     # bool[N] control flags (truth tables)
    arrays: dict = field(default_factory=dict)

    @property
    def w32(self):
        return (self.n + 31) // 32


def _split_node(node):
    """(mask, table, name, is_control) from a 4- or 5-tuple."""
    if len(node) == 5:
        mask, table, _i, name, ctrl = node
    elif len(node) == 4:
        mask, table, name, ctrl = node
    else:
        raise ValueError("PBN_data entries are (input_mask, truth_table, [i,] name, is_control)")
    return np.asarray(mask, dtype=bool), np.asarray(table, dtype=np.float64), name, bool(ctrl)


def compile_pbn_data(pbn_data) -> NetworkSpec:
    n = len(pbn_data)
    if n == 0:
        raise ValueError("empty PBN_data")
    in_off, ins, tab_off, prob, names, control = [0], [], [0], [], [], []
    for i, node in enumerate(pbn_data):
        mask, table, name, ctrl = _split_node(node)
        if mask.shape != (n,):
            raise ValueError(f"input mask of node {i} has length {mask.shape}, expected {n}")
        idx = np.nonzero(mask)[0]
        if table.size != 2 ** len(idx):
            raise ValueError(f"truth table of node {i} has {table.size} entries for {len(idx)} inputs")
        ins += idx.tolist()           # ascending node index; first masked node = MSB (common/node.py:31-32)
        in_off.append(len(ins))
        prob += table.reshape(-1).tolist()  # C order == .item(tuple(bits))
        tab_off.append(len(prob))
        names.append(str(name) if name is not None else f"G{i}")
        control.append(ctrl)
    arrays = dict(tt_in_off=np.array(in_off, np.int32), tt_in=np.array(ins + [0], np.int32),
                  tt_tab_off=np.array(tab_off, np.int32), tt_prob=np.array(prob, np.float64))
    return NetworkSpec(abi.NET_TT, n, 1, names, control=np.array(control, bool), arrays=arrays)


def predictor_lut16(A) -> int:
    """LUT bit (x0<<3 | x1<<2 | x2<<1 | x_self) = 0 iff np.matmul(X.T, A) < 0.  (bittner/base.py:100-118)."""
    A = np.asarray(A, dtype=np.float64).reshape(4, 1)
    lut = 0
    for idx in range(16):
        X = np.ones((4, 1))
        X[0], X[1], X[2], X[3] = (idx >> 3) & 1, (idx >> 2) & 1, (idx >> 1) & 1, idx & 1
        if not (np.matmul(X.T, A) < 0.0):
            lut |= 1 << idx
    return lut


def compile_predictor_sets(predictor_sets, node_ids, names=None) -> NetworkSpec:
    n = len(node_ids)
    if len(predictor_sets) != n:
        raise ValueError(f"{len(predictor_sets)} predictor sets for {n} node ids")
    where = {int(g): i for i, g in enumerate(node_ids)}
    off, ins, luts, cums, sums = [0], [], [], [], []
    for i in range(n):
        total, running = 0, None
        for cod, A, input_ids in np.asarray(predictor_sets[i], dtype=object).T:
            if cod is None:
                continue  # empty slot of the top-F buffer (gen/predictor_sets.py:80-102)
            total += cod
            running = cod if running is None else running + cod  # Node.add_predictors, base.py:36-42
            ins += [where[int(g)] for g in input_ids] + [i]   # own value sits where the bias would be (base.py:104)
            luts.append(predictor_lut16(A))
            cums.append(running)
        if running is None:
            raise ValueError(f"node {i} has no predictors")
        sums.append(total)
        off.append(len(luts))
    arrays = dict(pr_off=np.array(off, np.int32), pr_in=np.array(ins, np.int32), pr_lut=np.array(luts, np.uint16),
                  pr_cum=np.array(cums, np.float64), pr_codsum=np.array(sums, np.float64))
    return NetworkSpec(abi.NET_PRED, n, 0, names or [f"G{g}" for g in node_ids], ids=[int(g) for g in node_ids], arrays=arrays)


def bittner_available():
    return sorted(json.load(open(DATA_DIR / "node_ids.json")).keys())


def load_bittner(name: str) -> NetworkSpec:
    """One of the shipped predictor-set pickles, e.g. '100_5_kmeans' (node order recorded in data/node_ids.json)."""
    meta = json.load(open(DATA_DIR / "node_ids.json"))
    if name not in meta:
        raise FileNotFoundError(f"no shipped predictor set '{name}'; available: {sorted(meta)}")
    with open(DATA_DIR / f"predictor_sets_{name}.pkl", "rb") as f:
        sets = pickle.load(f)
    return compile_predictor_sets(sets, meta[name]["node_ids"])


def compile_cubes(n, attractors=(), targets=()):
    """-> (cube int8 [C][N] with 2 for '*', att_off int32 [A+1], tgt_first, n_tgt)."""
    rows, off = [], [0]
    for att in attractors:
        for cube in att:
            if len(cube) != n:
                raise ValueError(f"attractor state of length {len(cube)} for a network of {n} nodes")
            rows.append([2 if (isinstance(v, str) and v == "*") else int(v) for v in cube])
        off.append(len(rows))
    tgt_first = len(rows)
    for t in targets:
        if len(t) != n:
            raise ValueError(f"target state of length {len(t)} for a network of {n} nodes")
        rows.append([int(v) for v in t])
    cube = np.array(rows, np.int8).reshape(-1, n) if rows else np.zeros((1, n), np.int8)
    return cube, np.array(off, np.int32), tgt_first, len(rows) - tgt_first
