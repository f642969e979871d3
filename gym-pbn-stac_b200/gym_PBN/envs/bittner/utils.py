"""Bittner network construction from the SHIPPED predictor-set pickles (reference: gym_PBN/envs/bittner/utils.py).

The reference's `spawn` re-derives the network from genedata.xls (read -> pad ids -> binarise -> fit all gene triples)
and caches the fit as a pickle.  That offline pipeline is outside the hot path;
the pickles it produced are shipped, so `spawn` here resolves (total_genes, n_predictors, bin_method) to a shipped
set and compiles it for the GPU.
"""
from pathlib import Path

from gym_PBN.b200 import compiler

from . import base

DATA = Path(__file__).resolve().parent / "data"

# requested gene count -> shipped set (the registered envs ask for `{N}_3_median`, which upstream never shipped)
SHIPPED = {28: "28_15_median", 70: "70_5_kmeans", 100: "100_5_kmeans", 150: "150_5_kmeans", 200: "200_5_kmeans"}


def pad_ids(current_ids, pad_to, id_pool):
    """Append ids from the pool, in pool order, until `pad_to` are present (bittner/utils.py:42-51)."""
    new_ids = list(current_ids)
    for _id in id_pool:
        if _id not in new_ids:
            new_ids.append(_id)
            if len(new_ids) == pad_to:
                break
    return new_ids


def resolve_set(total_genes, n_predictors=None, bin_method=None):
    exact = f"{total_genes}_{n_predictors}_{bin_method}"
    if exact in compiler.bittner_available():
        return exact
    if total_genes in SHIPPED:
        return SHIPPED[total_genes]
    raise FileNotFoundError(
        f"no shipped predictor set for {total_genes} genes (available: {compiler.bittner_available()}); fitting new "
        "predictor sets from genedata.xls is the reference's offline pipeline and is not part of this build"
    )


def spawn(file=None, total_genes=None, include_ids=None, bin_method="median", n_predictors=5, predictor_sets_path=DATA,
          device=None, seed=None):
    name = resolve_set(total_genes, n_predictors, bin_method)
    return base.Graph.from_spec(compiler.load_bittner(name), device=device, seed=seed)
