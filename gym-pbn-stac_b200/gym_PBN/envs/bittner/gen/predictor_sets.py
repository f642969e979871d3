"""Predictor-set fitter (reference: gym_PBN/envs/bittner/gen/predictor_sets.py).

For every gene the reference scores every triple of other genes by the coefficient of determination of a rounded
least-squares fit (gen_COD, :105-124) and keeps the best `n_predictors` in a small buffer (add_to_buff, :80-102).
Here the scan over all candidates runs on the GPU (`pbn_fit_scan_host`, csrc/pbn_fit.cu: one 32-bit mask per gene row,
exact integer least squares, per-thread top lists merged per block and per gene); the host only
  * tabulates, per target row, the COD of "k misclassified samples" with the reference's own float expressions,
  * settles the few candidates whose rounding the reference leaves to float noise (a fitted value of exactly 1/2),
  * recomputes (COD, A) of the winners with the reference's float expressions (pinv), because A is stored in the
    predictor set and its sign pattern decides the network's lookup tables (base.py:110-118).

Buffer semantics reproduced (add_to_buff): slots 0..n-2 are the best n-1 candidates, earlier visits winning ties (strict
`<`); the loop bound `n_predictors - 1` (:85) means the LAST slot is never written directly — it holds whatever the most
recent insertion pushed out of slot n-2.  That is the (n-1)-th best among the candidates visited before the most recent
insertion, or None if fewer than n-1 were (why node 7 of the shipped 100/200-gene sets reads [1, 1, 1, 1, None]).
"""
import ctypes as C
import pickle
from pathlib import Path

import numpy as np

from gym_PBN.b200 import abi

RANK_MAX = 1023
ARR_BITS = 52
ARR_MASK = (1 << ARR_BITS) - 1
KEY_NONE = 0xFFFFFFFFFFFFFFFF


# ------------------------------------------------------------------------------------- reference float expressions
def null_error(y):
    """e_null of gen_COD (:117-118): MSE of the constant predictor round(mean(y)) + 1e-8."""
    y = np.asarray(y, dtype=np.float64).reshape(-1, 1)
    ones = np.ones(y.shape, dtype=np.float64)
    y_null = ((ones * np.mean(y)).round()).astype(int) + 10 ** -8
    return _mse(y_null, y)


def _mse(pred, y):
    # sklearn.metrics.mean_squared_error for one output: mean over samples, then mean over the single output
    return np.average(np.average((y - pred) ** 2, axis=0))


def cod_of_errors(e_null, k, n_samples):
    """COD when k of the n_samples squared errors are 1 (the rounded fit is 0/1): (e_null - k/n) / e_null, floored (:121-123)."""
    e = np.float64(k) / np.float64(n_samples)
    cod = (e_null - e) / e_null
    return 10 ** -8 if cod < 0 else cod


def gen_cod(x, y):
    """(COD, A, squared-error sum) of one candidate with the reference's float operations (gen_COD, :105-124).
    x int [S][k], y int [S]."""
    y = np.asarray(y, dtype=np.float64).reshape(-1, 1)
    ones = np.ones(y.shape, dtype=np.float64)
    X = np.append(np.asarray(x), ones, axis=1).astype(np.float64)
    R = np.dot(X.T, X)
    A = np.dot(np.linalg.pinv(R), np.dot(X.T, y))
    y_pred = np.dot(X, A).round()
    e_null = null_error(y)
    e = _mse(y_pred, y)
    cod = (e_null - e) / e_null
    if cod < 0:
        cod = 10 ** -8
    return cod, A, int(round(float(((y - y_pred) ** 2).sum())))


# ------------------------------------------------------------------------------------------------- gene table
class GeneTable:
    """Binarised rows grouped by gene id in order of first appearance (a gene id may own several rows: the spreadsheet
    lists some clones more than once and drop_duplicates only removes identical ones, bittner/utils.py:70-72)."""

    def __init__(self, ids, values):
        ids = np.asarray(ids)
        values = np.asarray(values)
        if values.ndim != 2 or len(ids) != values.shape[0]:
            raise ValueError("values must be [rows][samples] with one id per row")
        if not np.isin(values, (0, 1)).all():
            raise ValueError("the fitter expects binarised expression data")
        self.n_samples = values.shape[1]
        self.genes = list(dict.fromkeys(ids.tolist()))
        order = np.concatenate([np.nonzero(ids == g)[0] for g in self.genes]) if len(ids) else np.zeros(0, int)
        self.values = values[order].astype(np.int64)
        counts = [int((ids == g).sum()) for g in self.genes]
        self.row_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        self.row_gene = np.repeat(np.arange(len(self.genes)), counts)
        if self.n_samples <= 32:
            self.masks = (self.values.astype(np.uint64) << np.arange(self.n_samples, dtype=np.uint64)).sum(axis=1).astype(np.uint32)

    @classmethod
    def from_frame(cls, gene_data):
        data = gene_data.drop("Name", axis=1) if "Name" in getattr(gene_data, "columns", ()) else gene_data
        return cls(np.asarray(data.index), data.to_numpy())

    def rows_of(self, g):
        return range(self.row_off[g], self.row_off[g + 1])


def _rank_tables(table):
    """cod[R][S+1] float64 and rank[R][S+1] uint16: rank = class of the COD among all (row, k) of the SAME gene."""
    S = table.n_samples
    cod = np.empty((len(table.values), S + 1))
    for r, y in enumerate(table.values):
        e_null = null_error(y)
        for k in range(S + 1):
            cod[r, k] = cod_of_errors(e_null, k, S)
    rank = np.zeros(cod.shape, np.uint16)
    for g in range(len(table.genes)):
        rows = list(table.rows_of(g))
        classes = np.unique(cod[rows])[::-1]  # descending COD
        if len(classes) > RANK_MAX:
            raise ValueError("too many distinct COD values for one gene")
        rank[rows] = np.searchsorted(-classes, -cod[rows]).astype(np.uint16)
    return cod, rank


# --------------------------------------------------------------------------------------------------- key helpers
def decode_key(key):
    key = int(key)
    return dict(rank=key >> ARR_BITS, a=(key >> 40) & 0xFFF, b=(key >> 28) & 0xFFF, c=(key >> 16) & 0xFFF,
                y=(key >> 12) & 0xF, sc=key & 0xFFF)


def _candidate(table, g, key):
    """(x [S][3], y [S], input gene indices) of the candidate a key names."""
    d = decode_key(key)
    rem = [i for i in range(len(table.genes)) if i != g]
    ga, gb, gc = rem[d["a"]], rem[d["b"]], rem[d["c"]]
    na, nb, nc = (table.row_off[i + 1] - table.row_off[i] for i in (ga, gb, gc))
    sc = d["sc"]
    ia, ib, ic = sc // (nb * nc), (sc // nc) % nb, sc % nc
    x = np.stack([table.values[table.row_off[ga] + ia], table.values[table.row_off[gb] + ib],
                  table.values[table.row_off[gc] + ic]], axis=1)
    y_row = table.row_off[g] + d["y"]
    return x, table.values[y_row], (ga, gb, gc), y_row


def _scan(table, rank, top_l, key_gt=None, arr_lt=None, tie_le=None, tie_cap=1 << 16):
    """One device scan -> (sorted non-tie keys per gene, [(key_with_best_case_rank, gene)], kernel ms)."""
    G = len(table.genes)
    lib = abi.lib()
    desc = abi.PbnFitDesc(G, table.n_samples, table.row_off.ctypes.data, table.masks.ctypes.data, rank.ctypes.data)
    top = np.empty((G, top_l), np.uint64)
    ms = C.c_float(0)
    as_ptr = lambda a: None if a is None else a.ctypes.data  # noqa: E731
    while True:
        ties = np.zeros((tie_cap, 2), np.uint64) if tie_le is not None else None
        n_ties = C.c_int64(0)
        abi.check(lib.pbn_fit_scan_host(C.byref(desc), top_l, as_ptr(key_gt), as_ptr(arr_lt), as_ptr(tie_le),
                                        top.ctypes.data, as_ptr(ties), tie_cap, C.byref(n_ties), C.byref(ms)))
        if tie_le is None or n_ties.value <= tie_cap:
            break
        tie_cap = int(n_ties.value) + 1024  # rare: rerun with room for every reported candidate
    keys = []
    for g in range(G):
        keys.append([int(v) for v in top[g] if v != KEY_NONE])
    tie_list = [] if ties is None else [(int(k), int(g)) for k, g in ties[: n_ties.value]]
    return keys, tie_list, ms.value


def _settle(table, rank, ties, n_samples):
    """Exact keys of the rounding-tie candidates, by running the reference's float fit on each (rank from its actual error)."""
    out = {}
    for key, g in ties:
        x, y, _genes, y_row = _candidate(table, g, key)
        _cod, _A, k = gen_cod(x, y)
        r = int(rank[y_row, min(k, n_samples)])
        out.setdefault(g, []).append((r << ARR_BITS) | (key & ARR_MASK))
    return out


def fit_predictor_sets(table, n_predictors=5, stats=None):
    """list over genes of (3, n_predictors) object arrays (COD, A (4,1), input gene ids) — the reference's pickle payload."""
    G, S = len(table.genes), table.n_samples
    if S > 32:
        raise ValueError("the device fitter packs a gene row into one 32-bit mask: at most 32 samples")
    if n_predictors < 2 or n_predictors > 17:
        raise ValueError("n_predictors must be in 2..17")
    L = n_predictors - 1
    _cod, rank = _rank_tables(table)
    ms_total, n_settled = 0.0, 0

    def best_with_ties(top_l, key_gt=None, arr_lt=None, exclude=None):
        nonlocal ms_total, n_settled
        keys, _t, ms = _scan(table, rank, top_l, key_gt, arr_lt)
        ms_total += ms
        bound = np.array([(k[top_l - 1] >> ARR_BITS) if len(k) >= top_l else RANK_MAX for k in keys], np.uint16)
        _k, ties, ms = _scan(table, rank, top_l, key_gt, arr_lt, tie_le=bound)
        ms_total += ms
        n_settled += len(ties)
        settled = _settle(table, rank, ties, S)
        merged = []
        for g in range(G):
            extra = [k for k in settled.get(g, []) if (key_gt is None or k > int(key_gt[g])) and (exclude is None or k not in exclude[g])]
            merged.append(sorted(set(keys[g]) | set(extra))[:top_l])
        return merged

    top = best_with_ties(L)
    # last slot: the best candidate outside the final top list among those visited before the most recent insertion
    key_gt = np.array([k[-1] if len(k) == L else KEY_NONE for k in top], np.uint64)
    arr_lt = np.array([max(v & ARR_MASK for v in k) if len(k) == L else 0 for k in top], np.uint64)
    last = best_with_ties(1, key_gt, arr_lt, exclude=[set(k) for k in top])

    genes = np.asarray(table.genes)
    out = []
    for g in range(G):
        buff = np.empty((3, n_predictors), dtype=object)
        remaining = np.array([genes[i] for i in range(G) if i != g])
        entries = list(top[g])
        slots = list(range(len(entries)))
        if last[g]:
            entries.append(last[g][0])
            slots.append(n_predictors - 1)
        for slot, key in zip(slots, entries):
            x, y, _genes, _row = _candidate(table, g, key)
            c, A, _k = gen_cod(x, y)
            d = decode_key(key)
            buff[0, slot], buff[1, slot], buff[2, slot] = c, A, remaining[[d["a"], d["b"], d["c"]]]
        out.append(buff)
    if stats is not None:
        n_cand = 0
        for g in range(G):
            n = np.array([table.row_off[i + 1] - table.row_off[i] for i in range(G) if i != g], dtype=np.float64)
            e1, e2, e3 = n.sum(), (n ** 2).sum(), (n ** 3).sum()
            n_cand += (e1 ** 3 - 3 * e1 * e2 + 2 * e3) / 6 * (table.row_off[g + 1] - table.row_off[g])
        stats.update(kernel_ms=ms_total, scans=4, candidates=int(round(n_cand)), settled_on_host=n_settled)
    return out


def generate_predictor_sets(gene_data, k=3, n_predictors=5, savepath="predictor_sets.pkl", stats=None):
    """Reference signature (predictor_sets.py:13-38): load `savepath` if it exists, else fit on the GPU and save."""
    if Path(savepath).exists():
        with open(savepath, "rb") as f:
            return pickle.load(f)
    if k != 3:
        raise NotImplementedError("the device fitter scores gene TRIPLES (k = 3), the only arity the reference's networks use (base.py:100-104)")
    table = gene_data if isinstance(gene_data, GeneTable) else GeneTable.from_frame(gene_data)
    predictor_sets = fit_predictor_sets(table, n_predictors, stats)
    with open(savepath, "wb") as f:
        pickle.dump(predictor_sets, f)
    return predictor_sets
