"""Derives the node-ID order of each shipped predictor-set pickle (run once, output committed).

The pickles were produced by upstream `spawn` (bittner/utils.py:54-91) with the UNSORTED pad_ids
order (golden list tests/test_bittner.py:27); this fork's sorted order (utils.py:68) gives
self-inclusion violations, so we keep the unsorted order.  For the 150/200 sets
`drop_duplicates` (utils.py:72) removed one padded gene; candidates that keep every predictor
input resolvable and no self-inclusion are reported, and position 135 (gene 295483) is recorded.
"""
import json
import pickle
from pathlib import Path

DATA = Path(__file__).resolve().parent.parent / "gym-pbn-stac_b200" / "gym_PBN" / "envs" / "bittner" / "data"
INC7 = [234237, 324901, 759948, 25485, 266361, 108208, 130057]
INC28 = [234237, 324901, 759948, 25485, 324700, 43129, 266361, 108208, 40764, 130057, 39781, 49665,
         39159, 23185, 417218, 31251, 343072, 142076, 128100, 376725, 112500, 241530, 44563, 36950,
         812276, 51018, 306013, 418105]


def pad(cur, n, pool):
    new = list(cur)
    for i in pool:
        if i not in new:
            new.append(i)
            if len(new) == n:
                break
    return new


def violations(ids, sets):
    self_inc = missing = 0
    for i, x in enumerate(sets):
        for cod, _A, inp in x.T:
            if cod is None:
                continue
            for g in inp:
                self_inc += int(g == ids[i])
                missing += int(g not in ids)
    return self_inc, missing


def main():
    weights = json.load(open(DATA / "weighted_gene_ids.json"))
    out = {}
    for name, n, base in [("28_15_median", 28, INC28), ("70_5_kmeans", 70, INC7), ("70_5_kmeans-log", 70, INC7),
                          ("100_5_kmeans", 100, INC7), ("150_5_kmeans", 150, INC7), ("200_5_kmeans", 200, INC7)]:
        sets = pickle.load(open(DATA / f"predictor_sets_{name}.pkl", "rb"))
        ids = pad(base, n, weights) if n != len(base) else list(base)
        dropped = None
        if len(ids) != len(sets):
            cands = [d for d in range(len(ids)) if violations(ids[:d] + ids[d + 1:], sets) == (0, 0)]
            assert 135 in cands, cands
            dropped = {"position": 135, "gene": ids[135], "candidates": cands}
            ids = ids[:135] + ids[136:]
        assert violations(ids, sets) == (0, 0), name
        out[name] = {"node_ids": [int(i) for i in ids], "dropped": dropped}
        print(name, len(ids), dropped)
    json.dump(out, open(DATA / "node_ids.json", "w"))


if __name__ == "__main__":
    main()
