"""TEST INFRASTRUCTURE — CPU restatement of the reference's predictor-set fitter and gene-data preparation.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product
(gym_PBN.envs.bittner.gen) never does.

Follows, one candidate at a time and with the same NumPy float operations:
  * gym_PBN/envs/bittner/gen/predictor_sets.py:41-78   _gen_predictor_sets_gene (visiting order of candidates)
  * gym_PBN/envs/bittner/gen/predictor_sets.py:80-102  add_to_buff (insertion rule, loop bound n_predictors - 1)
  * gym_PBN/envs/bittner/gen/predictor_sets.py:105-124 gen_COD (pinv fit, rounding, 1e-8 null offset and floor)
  * gym_PBN/envs/bittner/gen/binarise.py:10-35,179-199 binarise (median / average / the legacy "kmeans")
  * gym_PBN/envs/bittner/utils.py:42-51,54-72          pad_ids, the trim / binarise / drop_duplicates steps of spawn

PINNED: the shipped predictor_sets_28_15_median.pkl was produced by the reference's own fitter from genedata.xls with
the ids of tests/test_bittner.py:83; tests/test_oracle_fit.py checks that this restatement reproduces its CODs and input
ids exactly (coefficients A to 1e-9: they come out of LAPACK's SVD and differ in the last bits between builds), and that
the kmeans rule reproduces /root/reference/binned_test.csv (committed as tests/golden/fit_binned_70_kmeans.npz).
"""
import itertools

import numpy as np


def mse(pred, y):
    return np.average(np.average((y - pred) ** 2, axis=0))


def gen_cod(x, y):
    """x [S][k] ints, y [S][1] ints -> (COD, A)."""
    ones = np.ones(y.shape, dtype=np.float64)
    X = np.append(x, ones, axis=1).astype(np.float64)
    Y = y.astype(np.float64)
    gram = np.dot(X.T, X)
    A = np.dot(np.linalg.pinv(gram), np.dot(X.T, Y))
    fitted = np.dot(X, A).round()
    null = ((ones * np.mean(Y)).round()).astype(int) + 10 ** -8
    e_null, e = mse(null, Y), mse(fitted, Y)
    cod = (e_null - e) / e_null
    return (10 ** -8 if cod < 0 else cod), A


def push(buff, entry):
    """add_to_buff: only slots 0..n-2 are compared against; the last slot receives what a shift pushes out."""
    n = buff.shape[1]
    i = 0
    while i < n - 1:
        if buff[0, i] is None:
            buff[:, i] = entry
            return
        if buff[0, i] < entry[0]:
            carry = buff[:, i].copy()
            buff[:, i] = entry
            for j in range(i + 1, n):
                carry, buff[:, j] = buff[:, j].copy(), carry
            return
        i += 1


def fit_gene(ids, values, gene, n_predictors, k=3):
    """(3, n_predictors) object buffer of one gene.  ids [R] (repeats allowed), values [R][S] in {0,1}."""
    ids = np.asarray(ids)
    genes = list(dict.fromkeys(ids.tolist()))
    remaining = np.array([g for g in genes if g != gene])
    target_rows = values[ids == gene]
    rows_of = [values[ids == g] for g in remaining]
    buff = np.empty((3, n_predictors), dtype=object)
    for comb in itertools.combinations(range(len(remaining)), k):
        xs = [np.array(rows).T for rows in itertools.product(*[rows_of[c] for c in comb])]
        for y in target_rows:
            for x in xs:
                cod, A = gen_cod(x, y[:, None])
                entry = np.empty(3, dtype=object)
                entry[0], entry[1], entry[2] = cod, A, remaining[list(comb)]
                push(buff, entry)
    return buff


def fit_all(ids, values, n_predictors, k=3):
    return [fit_gene(ids, values, g, n_predictors, k) for g in dict.fromkeys(np.asarray(ids).tolist())]


# ------------------------------------------------------------------------------------------- data preparation
def thresholds(x, method):
    x = np.asarray(x, dtype=np.float64)
    if method == "median":
        return np.nanmedian(x, axis=1)
    if method == "average":
        return np.nanmean(x, axis=1)
    if method == "kmeans":
        # KMeansLegacyV2 with its aliased cluster lists: every run returns (max + min) / 2 of log1p(x)
        lx = np.log1p(x)
        return np.expm1((lx.max(axis=1) + lx.min(axis=1)) / 2)
    raise Exception(f'"{method}" is not an implemented binarisation method.')


def binarise(x, method):
    with np.errstate(invalid="ignore"):
        return (np.asarray(x, dtype=np.float64) > thresholds(x, method)[:, None]).astype(np.int64)


def pad_ids(current, pad_to, pool):
    out = list(current)
    for i in pool:
        if i not in out:
            out.append(i)
            if len(out) == pad_to:
                break
    return out


def prepare(all_ids, all_names, all_ratios, weight_ids, total_genes, include_ids, bin_method):
    """(ids [R], names [R], binarised [R][S]) after pad -> sort -> trim -> binarise -> drop duplicate rows."""
    include_ids = list(include_ids)
    if total_genes != len(include_ids):
        include_ids = sorted(pad_ids(include_ids, total_genes, weight_ids))
    all_ids = np.asarray(all_ids)
    sel = np.concatenate([np.nonzero(all_ids == g)[0] for g in include_ids])
    values = binarise(np.asarray(all_ratios)[sel], bin_method)
    seen, keep = set(), []
    for j, r in enumerate(sel):
        sig = (all_names[r],) + tuple(values[j])
        if sig not in seen:
            seen.add(sig)
            keep.append(j)
    return all_ids[sel][keep], [all_names[r] for r in sel[keep]], values[keep]
