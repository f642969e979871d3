"""Bittner network construction (reference: gym_PBN/envs/bittner/utils.py).

`spawn` is the reference's pipeline: read genedata.xls -> pad the gene list from the weighted list -> sort -> trim ->
binarise -> drop duplicate rows -> fit the predictor sets (cached as a pickle) -> build the graph.  The fit is the GPU
scan of gen/predictor_sets.py (csrc/pbn_fit.cu); reading, binarising and trimming are small host steps.

One deliberate addition: upstream ships fitted sets for 28/70/100/150/200 genes (kmeans binarisation, 5 or 15
predictors) while its registered envs ask for `{N}_3_median`, which it never shipped and would fit for hours on first
use.  With `predictor_set=None` (default) `spawn` uses the exact cached/shipped file when there is one, else the
shipped set of that size when there is one, else fits on the GPU.  `predictor_set="fit"` always takes the reference's
route (exact cache, else fit); `predictor_set="100_5_kmeans"` names a shipped set.
"""
import pickle
from pathlib import Path

import numpy as np

from gym_PBN.b200 import compiler

from . import base
from .gen.binarise import binarise_values
from .gen import predictor_sets as _fit
from .gen import xls as _xls

DATA = Path(__file__).resolve().parent / "data"

# requested gene count -> shipped set (the registered envs ask for `{N}_3_median`, which upstream never shipped)
SHIPPED = {28: "28_15_median", 70: "70_5_kmeans", 100: "100_5_kmeans", 150: "150_5_kmeans", 200: "200_5_kmeans"}


def extract_gene_data(file):
    """(gene_data DataFrame indexed by "ID" with columns Name, T1..T31; weight_ids Series) — bittner/utils.py:10-39."""
    import pandas as pd

    ids, names, ratios, weight_ids = _xls.read_gene_data(file)
    gene_data = pd.DataFrame(ratios, columns=[f"T{i + 1}" for i in range(ratios.shape[1])])
    gene_data.insert(0, "Name", names)
    gene_data.index = pd.Index(ids, name="ID")
    return gene_data, pd.Series(weight_ids, name="ID")


def pad_ids(current_ids, pad_to, id_pool):
    """Append ids from the pool, in pool order, until `pad_to` are present (bittner/utils.py:42-51)."""
    new_ids = list(current_ids)
    for _id in id_pool:
        if _id not in new_ids:
            new_ids.append(_id)
            if len(new_ids) == pad_to:
                break
    return new_ids


def prepare_gene_table(file, total_genes, include_ids, bin_method):
    """(include_ids after padding/sorting, ids [R], names [R], binarised int64 [R][S]) — the trim / binarise /
    drop_duplicates steps of spawn (bittner/utils.py:62-72) on plain arrays."""
    ids, names, ratios, weight_ids = _xls.read_gene_data(file)
    include_ids = [int(i) for i in include_ids]
    if total_genes != len(include_ids):
        include_ids = sorted(pad_ids(include_ids, total_genes, weight_ids))  # sorted "for interoperability with cabean"
    missing = [g for g in include_ids if not (ids == g).any()]
    if missing:
        raise KeyError(f"gene ids not in the spreadsheet: {missing}")
    sel = np.concatenate([np.nonzero(ids == g)[0] for g in include_ids])
    values = binarise_values(ratios[sel], bin_method)
    seen, keep = set(), []
    for j, r in enumerate(sel):  # DataFrame.drop_duplicates(): same name and same binarised row
        sig = (names[r],) + tuple(values[j].tolist())
        if sig not in seen:
            seen.add(sig)
            keep.append(j)
    return include_ids, ids[sel][keep], [names[r] for r in sel[keep]], values[keep]


def fitted_spec(file, total_genes, include_ids, bin_method, n_predictors, predictor_sets_path=DATA, stats=None):
    """NetworkSpec of the reference pipeline: exact cache file if present, else a GPU fit (cached when the directory is
    writable).  File name as the reference builds it (bittner/utils.py:78): predictor_sets_{len(ids)}_{F}_{method}.pkl."""
    include_ids, ids, names, values = prepare_gene_table(file, total_genes, include_ids, bin_method)
    table = _fit.GeneTable(ids, values)
    savepath = Path(predictor_sets_path) / f"predictor_sets_{len(include_ids)}_{n_predictors}_{bin_method}.pkl"
    if savepath.exists():
        with open(savepath, "rb") as f:
            sets = pickle.load(f)
    else:
        sets = _fit.fit_predictor_sets(table, n_predictors, stats)
        try:
            with open(savepath, "wb") as f:
                pickle.dump(sets, f)
        except OSError:
            pass
    first_name = {}
    for g, nm in zip(ids.tolist(), names):
        first_name.setdefault(g, nm)
    return compiler.compile_predictor_sets(sets, table.genes, [str(first_name[g]) for g in table.genes])


def resolve_set(total_genes, n_predictors=None, bin_method=None):
    """Name of a shipped set for this request, or None."""
    exact = f"{total_genes}_{n_predictors}_{bin_method}"
    if exact in compiler.bittner_available():
        return exact
    return SHIPPED.get(total_genes)


def spawn(file=None, total_genes=None, include_ids=None, bin_method="median", n_predictors=5, predictor_sets_path=DATA,
          device=None, seed=None, predictor_set=None):
    if predictor_set not in (None, "fit"):
        return base.Graph.from_spec(compiler.load_bittner(predictor_set), device=device, seed=seed)
    name = None if predictor_set == "fit" else resolve_set(total_genes, n_predictors, bin_method)
    if name is not None:
        return base.Graph.from_spec(compiler.load_bittner(name), device=device, seed=seed)
    if include_ids is None:
        raise FileNotFoundError(f"no shipped predictor set for {total_genes} genes and no include_ids to fit one from "
                                f"(available: {compiler.bittner_available()})")
    spec = fitted_spec(file or DATA / "genedata.xls", total_genes, include_ids, bin_method, n_predictors, predictor_sets_path)
    return base.Graph.from_spec(spec, device=device, seed=seed)
