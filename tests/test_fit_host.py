"""CPU: gene-data preparation and the host side of the predictor-set fitter.

The first five tests mirror the reference's own tests/test_bittner.py (same constants).  The GPU scan itself is covered by
tests/test_gpu_fit.py; here the host orchestration runs on a NumPy stand-in of the scan (tests/fit_emul.py) that scores
candidates with the library's own per-candidate solver executed on the host.
"""
import json
from fractions import Fraction
from math import floor

import numpy as np
import pytest

from gym_PBN.b200 import abi
from gym_PBN.envs.bittner import utils
from gym_PBN.envs.bittner.gen import binarise, generate_predictor_sets  # noqa: F401  (reference import surface)
from gym_PBN.envs.bittner.gen import predictor_sets as ps
from gym_PBN.envs.bittner.gen.binarise import binarise_values

GENE_DATA_PATH = utils.DATA / "genedata.xls"
M7 = [234237, 324901, 759948, 25485, 266361, 108208, 130057]


@pytest.fixture(scope="module")
def extracted():
    return utils.extract_gene_data(GENE_DATA_PATH)


def test_data_extraction(extracted):  # tests/test_bittner.py:12-18
    gene_data, weight_ids = extracted
    assert gene_data.shape[0] == 8067
    assert gene_data.shape[1] == 32
    assert gene_data.index.name == "ID"
    assert len(weight_ids) == 276


def test_id_padding(extracted, golden_dir):  # tests/test_bittner.py:21-29 (golden list = first 70 padded ids)
    _, weight_ids = extracted
    new_ids = utils.pad_ids(M7, 70, weight_ids)
    meta = json.load(open(utils.DATA / "node_ids.json"))
    assert new_ids == meta["70_5_kmeans"]["node_ids"]  # the upstream 70-gene set was fitted in exactly this order
    assert new_ids == np.load(golden_dir / "fit_prepare.npz")["padded"].tolist()
    assert new_ids[:12] == [234237, 324901, 759948, 25485, 266361, 108208, 130057, 357278, 39781, 49665, 39159, 23185]
    assert len(set(utils.pad_ids(M7, 200, weight_ids))) == 200


def test_trimming_and_loc(extracted):  # tests/test_bittner.py:32-39,67-79
    gene_data, weight_ids = extracted
    new_ids = utils.pad_ids(M7, 70, weight_ids)
    trimmed = gene_data.loc[new_ids]
    assert trimmed.shape == (85, 32)
    first = []
    for i in trimmed.index:
        j = np.where(trimmed.index == i)[0][0]
        if j not in first:
            first.append(j)
    assert len(first) == len(new_ids)


def test_binarisation_matches_reference(extracted, golden_dir):  # tests/test_bittner.py:42-64 + recorded reference outputs
    gene_data, weight_ids = extracted
    z = np.load(golden_dir / "fit_prepare.npz")
    trimmed = gene_data.loc[utils.pad_ids(M7, 70, weight_ids)]
    assert np.array_equal(np.asarray(trimmed.index), z["ids"])
    assert np.array_equal(trimmed.drop("Name", axis=1).to_numpy(), z["ratios"])
    t_cols = [f"T{i}" for i in range(1, 32)]
    for method in ("median", "average", "kmeans"):
        np.random.seed(0)
        binned = binarise(trimmed, method)
        assert binned.shape[1] == 32 and binned.index.name == "ID" and list(binned.columns)[-1] == "Name"
        assert all(t == np.int64 for t in binned.dtypes[t_cols])
        assert np.array_equal(binned[t_cols].to_numpy(), z[method]), method
        assert np.random.rand(1)[0] == z[method + "_next_draw"][0]  # the legacy stream ends where the reference leaves it
        assert binned[t_cols].to_numpy().max() == 1 and binned[t_cols].to_numpy().min() == 0
    with pytest.raises(Exception, match="not an implemented"):
        binarise(trimmed, "otsu")
    assert np.array_equal(binarise_values(z["ratios"], "median"), z["median"])


def test_prepare_gene_table_orders_and_duplicates():
    meta = json.load(open(utils.DATA / "node_ids.json"))
    b28 = meta["28_15_median"]["node_ids"]
    inc, ids, names, values = utils.prepare_gene_table(GENE_DATA_PATH, 28, b28, "median")
    assert inc == b28 and len(ids) == 36 and values.shape == (36, 31) and len(names) == 36  # 37 rows, one exact duplicate
    table = ps.GeneTable(ids, values)
    assert table.genes == b28 and table.row_off[-1] == 36 and max(np.diff(table.row_off)) == 2
    assert all(int(table.masks[r]) == sum(int(v) << s for s, v in enumerate(table.values[r])) for r in range(36))
    inc, ids, _n, values = utils.prepare_gene_table(GENE_DATA_PATH, 10, M7, "median")
    assert inc == sorted(inc) and len(inc) == 10 and set(M7) <= set(inc)  # padded lists are sorted (bittner/utils.py:66)
    with pytest.raises(KeyError):
        utils.prepare_gene_table(GENE_DATA_PATH, 2, [234237, 1], "median")


# ------------------------------------------------------------------------------------------------- exact solver
def _exact(ma, mb, mc, my, S):
    cols = [[1] * S] + [[(m >> s) & 1 for s in range(S)] for m in (ma, mb, mc)]
    y = [(my >> s) & 1 for s in range(S)]
    dot = lambda u, v: sum(Fraction(a) * b for a, b in zip(u, v))  # noqa: E731
    ortho = []
    for c in cols:  # Gram-Schmidt over the rationals: projection onto the column space, whatever its rank
        v = [Fraction(x) for x in c]
        for o in ortho:
            f = dot(v, o) / dot(o, o)
            v = [a - f * b for a, b in zip(v, o)]
        if any(v):
            ortho.append(v)
    fit = [Fraction(0)] * S
    for o in ortho:
        f = dot(y, o) / dot(o, o)
        fit = [a + f * b for a, b in zip(fit, o)]
    # half-integer fitted values: both roundings, chosen per input pattern (all samples of a pattern share the fit)
    pats = {}
    for s in range(S):
        pats.setdefault((cols[1][s], cols[2][s], cols[3][s]), []).append(s)
    lo = hi = 0
    for ss in pats.values():
        f = fit[ss[0]]
        q = floor(f + Fraction(1, 2))
        e_q = sum((q - y[s]) ** 2 for s in ss)
        if (f + Fraction(1, 2)).denominator == 1:
            e_m = sum((q - 1 - y[s]) ** 2 for s in ss)
            lo, hi = lo + min(e_q, e_m), hi + max(e_q, e_m)
        else:
            lo, hi = lo + e_q, hi + e_q
    return lo, hi


@pytest.mark.parametrize("S", [31, 32, 12, 5, 1])
def test_exact_solver_against_rational_arithmetic(S):
    rng = np.random.default_rng(S)
    n, full = 250, (1 << S) - 1
    m = rng.integers(0, 1 << S, size=(n, 4), dtype=np.uint64).astype(np.uint32)
    m[0] = [0, 0, 0, full & 5]                      # constant inputs
    m[1] = [full, full, full, full & 3]
    m[2, 1] = m[2, 0]                               # identical inputs
    m[3, 1] = ~m[3, 0] & full                       # complementary inputs
    m[4, 3] = m[4, 0]                               # target equals an input
    m[5] = [m[5, 0]] * 4
    m[6, 2] = m[6, 0] & m[6, 1]
    m[7, 3] = 0                                     # constant target
    m[8, 3] = full
    lo, hi = np.zeros(n, np.int32), np.zeros(n, np.int32)
    abi.check(abi.lib().pbn_fit_eval_host(m.ctypes.data, n, S, lo.ctypes.data, hi.ctypes.data))
    for i in range(n):
        assert _exact(*(int(v) for v in m[i]), S) == (lo[i], hi[i]), (S, i, m[i])


def test_exact_solver_agrees_with_the_float_fit_away_from_ties():
    rng = np.random.default_rng(3)
    S, n = 31, 400
    m = rng.integers(0, 1 << S, size=(n, 4), dtype=np.uint64).astype(np.uint32)
    lo, hi = np.zeros(n, np.int32), np.zeros(n, np.int32)
    abi.check(abi.lib().pbn_fit_eval_host(m.ctypes.data, n, S, lo.ctypes.data, hi.ctypes.data))
    bits = lambda v: np.array([(int(v) >> s) & 1 for s in range(S)])  # noqa: E731
    for i in range(n):
        x = np.stack([bits(m[i, 0]), bits(m[i, 1]), bits(m[i, 2])], axis=1)
        _cod, _A, k = ps.gen_cod(x, bits(m[i, 3]))
        assert lo[i] <= k <= hi[i]
        if lo[i] == hi[i]:
            assert k == lo[i]


# ----------------------------------------------------------------------------------------- host orchestration
def _assert_same_sets(a, b):
    assert len(a) == len(b)
    for g, (x, y) in enumerate(zip(a, b)):
        assert x.shape == y.shape
        for s in range(x.shape[1]):
            assert (x[0, s] is None) == (y[0, s] is None), (g, s)
            if x[0, s] is not None:
                assert x[0, s] == y[0, s], (g, s, x[0, s], y[0, s])
                assert np.array_equal(x[2, s], y[2, s]), (g, s, x[2, s], y[2, s])
                assert np.array_equal(x[1, s], y[1, s]), (g, s)


def random_table(rng, G, S, max_rows=3, p_dup=0.3):
    ids, rows = [], []
    for g in range(G):
        for _ in range(1 + (rng.random() < p_dup) * int(rng.integers(1, max_rows))):
            ids.append(1000 + 7 * g)
            rows.append(rng.integers(0, 2, S))
    rows = np.array(rows)
    if G >= 4:
        rows[-1] = rows[0]              # a gene identical to another: COD 1 for every triple containing it
        rows[len(rows) // 2] = 1        # a constant gene (singular Gram matrix)
    return np.array(ids), rows


@pytest.mark.parametrize("G,S,F", [(6, 31, 3), (7, 8, 5), (8, 5, 4), (5, 32, 2), (4, 31, 3), (3, 31, 3), (9, 31, 17)])
def test_host_orchestration_matches_oracle(monkeypatch, G, S, F):
    import fit_emul
    import fit_oracle

    monkeypatch.setattr(ps, "_scan", fit_emul.scan)
    rng = np.random.default_rng(100 * G + S)
    ids, rows = random_table(rng, G, S)
    stats = {}
    ours = ps.fit_predictor_sets(ps.GeneTable(ids, rows), F, stats)
    _assert_same_sets(ours, fit_oracle.fit_all(ids, rows, F))
    if G == 3:
        assert all(v is None for b in ours for v in b.reshape(-1))  # no triple of OTHER genes exists
    if S <= 8 and G >= 7:
        assert stats["settled_on_host"] > 0  # few samples: half-integer fitted values do occur and are settled


def test_fit_argument_errors():
    lib = abi.lib()
    rank = np.zeros((2, 32), np.uint16)
    off = np.array([0, 1, 2], np.int32)
    rows = np.array([1, 2], np.uint32)
    top = np.zeros(64, np.uint64)

    def call(desc, top_l=2):
        return lib.pbn_fit_scan_host(desc, top_l, None, None, None, top.ctypes.data, None, 0, None, None)

    import ctypes as C
    good = abi.PbnFitDesc(2, 31, off.ctypes.data, rows.ctypes.data, rank.ctypes.data)
    for bad in (abi.PbnFitDesc(2, 33, off.ctypes.data, rows.ctypes.data, rank.ctypes.data),
                abi.PbnFitDesc(0, 31, off.ctypes.data, rows.ctypes.data, rank.ctypes.data),
                abi.PbnFitDesc(2, 1, off.ctypes.data, rows.ctypes.data, rank.ctypes.data)):  # mask bits beyond n_samples
        assert call(C.byref(bad)) == 1
    assert call(C.byref(good), top_l=17) == 1 and call(C.byref(good), top_l=0) == 1
    assert b"top_l" in lib.pbn_last_error()
    with pytest.raises(ValueError):
        ps.fit_predictor_sets(ps.GeneTable(np.arange(5), np.zeros((5, 40), int)), 3)
    with pytest.raises(ValueError):
        ps.GeneTable(np.arange(2), np.array([[0, 2], [1, 0]]))
    with pytest.raises(NotImplementedError):
        generate_predictor_sets(ps.GeneTable(np.arange(5), np.zeros((5, 4), int)), k=2, savepath="/nonexistent/x.pkl")
