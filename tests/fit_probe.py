"""TEST INFRASTRUCTURE (it runs the CPU oracle next to the product, so it lives under tests/).
GPU probe of the predictor-set fitter: fits the shipped configurations on the device, compares with the shipped
pickles (reference outputs) and with the CPU oracle on this machine, and prints kernel times.
  python tests/fit_probe.py [--oracle-genes 4] > gpurun_out/fit_probe.json"""
import argparse
import json
import pickle
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "gym-pbn-stac_b200"), str(ROOT / "oracle")]

from gym_PBN.envs.bittner import utils  # noqa: E402
from gym_PBN.envs.bittner.gen import predictor_sets as ps  # noqa: E402

B28 = [234237, 324901, 759948, 25485, 324700, 43129, 266361, 108208, 40764, 130057, 39781, 49665, 39159, 23185, 417218,
       31251, 343072, 142076, 128100, 376725, 112500, 241530, 44563, 36950, 812276, 51018, 306013, 418105]
M7 = [234237, 324901, 759948, 25485, 266361, 108208, 130057]


def compare(sets, ref, table):
    """#slots equal, #slots that differ, of which explained by rounding-tie candidates."""
    from gym_PBN.b200 import abi

    lib = abi.lib()

    def is_tie(g, idarr):
        import itertools
        rows = [list(table.rows_of(table.genes.index(int(i)))) for i in idarr]
        for y in table.rows_of(g):
            for ia, ib, ic in itertools.product(*rows):
                m = np.array([[table.masks[ia], table.masks[ib], table.masks[ic], table.masks[y]]], np.uint32)
                lo, hi = np.zeros(1, np.int32), np.zeros(1, np.int32)
                lib.pbn_fit_eval_host(m.ctypes.data, 1, table.n_samples, lo.ctypes.data, hi.ctypes.data)
                if lo[0] != hi[0]:
                    return True
        return False

    same = diff = genes_diff = genes_tie = 0
    cod_err = 0.0
    for g in range(len(ref)):
        bad = False
        for s in range(ref[g].shape[1]):
            a, b = sets[g][:, s], ref[g][:, s]
            if a[0] is None and b[0] is None:
                same += 1
            elif a[0] is not None and b[0] is not None and (np.asarray(a[2]) == np.asarray(b[2])).all():
                same += 1
                cod_err = max(cod_err, abs(a[0] - b[0]) / abs(b[0]))
            else:
                diff += 1
                bad = True
        if bad:
            genes_diff += 1
            cands = [sets[g][2, s] for s in range(ref[g].shape[1]) if sets[g][0, s] is not None]
            cands += [ref[g][2, s] for s in range(ref[g].shape[1]) if ref[g][0, s] is not None]
            genes_tie += any(is_tie(g, c) for c in cands)
    return dict(slots_same=same, slots_diff=diff, genes_diff=genes_diff, genes_diff_with_tie_candidate=genes_tie,
                max_rel_cod_err=cod_err)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--oracle-genes", type=int, default=4)
    ap.add_argument("--only", type=int, default=0, help="fit just the {N}_3_median configuration (profiling)")
    args = ap.parse_args()
    out = {}
    meta = json.load(open(utils.DATA / "node_ids.json"))

    def run(name, table, n_pred, ref=None, oracle_genes=0, ids=None, values=None):
        stats = {}
        t = time.time()
        sets = ps.fit_predictor_sets(table, n_pred, stats)
        rec = dict(genes=len(table.genes), rows=len(table.values), n_predictors=n_pred, wall_s=round(time.time() - t, 3), **stats)
        if ref is not None:
            rec["vs_shipped"] = compare(sets, ref, table)
        if oracle_genes:
            import fit_oracle as fo
            pick = list(range(0, len(table.genes), max(1, len(table.genes) // oracle_genes)))[:oracle_genes]
            t = time.time()
            ok = True
            for g in pick:
                o = fo.fit_gene(ids, values, table.genes[g], n_pred)
                for s in range(n_pred):
                    if (o[0, s] is None) != (sets[g][0, s] is None):
                        ok = False
                    elif o[0, s] is not None:
                        ok &= bool(o[0, s] == sets[g][0, s] and (o[2, s] == sets[g][2, s]).all() and (o[1, s] == sets[g][1, s]).all())
            rec["vs_oracle"] = dict(genes=pick, exact=ok, oracle_s=round(time.time() - t, 2))
        out[name] = rec
        print(name, json.dumps(rec), file=sys.stderr, flush=True)
        return sets

    if args.only:
        _i, ids, _n, values = utils.prepare_gene_table(utils.DATA / "genedata.xls", args.only, M7, "median")
        run(f"b{args.only}_3_median", ps.GeneTable(ids, values), 3)
        print(json.dumps(out))
        return
    _i, ids, _n, values = utils.prepare_gene_table(utils.DATA / "genedata.xls", 28, B28, "median")
    table = ps.GeneTable(ids, values)
    ref = pickle.load(open(utils.DATA / "predictor_sets_28_15_median.pkl", "rb"))
    run("b28_15_median", table, 15, ref, args.oracle_genes, ids, values)

    for n in (7, 10, 30, 50, 70, 100, 200):
        inc = sorted(M7) if n == 7 else M7
        _i, ids, _n, values = utils.prepare_gene_table(utils.DATA / "genedata.xls", n, inc, "median")
        run(f"b{n}_3_median", ps.GeneTable(ids, values), 3, None, 2 if n <= 30 else 0, ids, values)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
