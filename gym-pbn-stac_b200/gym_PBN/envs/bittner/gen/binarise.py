"""Binarisation of the expression ratios (reference: gym_PBN/envs/bittner/gen/binarise.py).

A gene row becomes 1 where its ratio is strictly above the row's threshold (binarise.py:33-35).  Thresholds:
  * "median" / "average" — the row's median / mean over its non-missing samples (pandas semantics, :12-13);
  * "kmeans" — what the reference's KMeansLegacyV2 actually returns (:179-199).  Its two "clusters" are one aliased
    list (`[[]] * n_clusters`, :93), so every run ends with both means equal to the overall mean and the "threshold
    between the clusters" (:103-111) is (max + min) / 2 of log1p(ratio): the mid-range in log space, independent of the
    random initialisation.  Every one of the n_init = 10 runs gives the same threshold, so the Gaussian-integral score
    (:124-150) cannot change the choice.  The reference still draws np.random.rand(2) per run and gene (:89); those
    draws are consumed here too so that the legacy global stream ends where the reference leaves it.
    The shipped kmeans predictor sets and the reference's binned_test.csv are reproduced bit for bit by this rule
    (tests/test_fit_host.py).
"""
import numpy as np

METHODS = ("average", "median", "kmeans")
KMEANS_N_INIT = 10  # KMeansLegacy.__init__ defaults, binarise.py:78-81


def thresholds(values, method, consume_draws=True):
    """float64 [R] thresholds for rows of `values` [R][S] (NaN = missing sample)."""
    x = np.asarray(values, dtype=np.float64)
    if method == "median":
        return np.nanmedian(x, axis=1)
    if method == "average":
        return np.nanmean(x, axis=1)
    if method == "kmeans":
        if np.isnan(x).any():
            raise ValueError("kmeans binarisation needs complete rows (the reference's KMeansLegacy has no missing-value handling)")
        x_log = np.log1p(x)
        out = np.empty(x.shape[0])
        for i in range(x.shape[0]):
            if consume_draws:
                for _ in range(KMEANS_N_INIT):
                    np.random.rand(2)
            out[i] = np.expm1((x_log[i].max() + x_log[i].min()) / 2)
        return out
    raise Exception(f'"{method}" is not an implemented binarisation method.')


def binarise_values(values, method, consume_draws=True):
    """int64 [R][S]: 1 where the ratio is strictly above the row's threshold."""
    x = np.asarray(values, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        return (x > thresholds(x, method, consume_draws)[:, None]).astype(np.int64)


def binarise(df, method):
    """DataFrame in, DataFrame out, as the reference: columns T1..Tn binarised, "Name" kept as the last column."""
    if method not in METHODS:
        raise Exception(f'"{method}" is not an implemented binarisation method.')
    names = df["Name"]
    data = df.drop("Name", axis=1)
    out = data.copy()
    out.iloc[:, :] = 0
    out = out.astype(np.int64)
    out.iloc[:, :] = binarise_values(data.to_numpy(dtype=np.float64), method)
    out["Name"] = names
    return out
