"""TEST INFRASTRUCTURE — a large steady-state-distribution estimate with the reference's LITERAL algorithm, for the tight
tolerance of tests/test_gpu_parity.py.

`_ssd_run` (gym_PBN/utils/eval.py:76-103) per chain and iteration: histogram the target genes, `np.random.rand(N) <
bit_flip_prob` -> flipNode, env.step(0) = one asynchronous update (randint node, uniform predictor pick).  This script runs
exactly that through the oracle's REPLAY path (float64 compares, N Bernoulli draws per iteration — not the geometric-skip
stream the product uses) fed by NumPy's generator: 32 independent batches x 8000 chains x 4000 iterations (the reference's
default chain length, utils/eval.py:23-24) = 1.024e9 iterations on Bittner-100 (100_5_kmeans, 7 target genes).  The oracle's
replay path itself is pinned to traces recorded from the unmodified reference (tests/golden/b100_ssd_replay.npz).

    python oracle/make_ssd_literal_golden.py   # ~3 min on 8 cores -> tests/golden/b100_ssd_literal.npz (per-batch histograms)
"""
import multiprocessing as mp
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
NAME, P, ITERS, BATCHES, CHAINS, CHUNK = "100_5_kmeans", 0.01, 4000, 32, 8000, 32
TGT = np.arange(7, dtype=np.int32)


def batch(b):
    import os

    os.environ["OMP_NUM_THREADS"] = "1"
    import oracle as orc

    sets, ids = orc.load_bittner(NAME)
    net = orc.net_from_predictor_sets(sets, ids)
    env = orc.Env(orc.ENV_TARGET, net.n)  # all-attracting: one update per iteration
    rng = np.random.default_rng([2024, b])
    hist = np.zeros(128, np.int64)
    for c0 in range(0, CHAINS, CHUNK):
        st = rng.integers(0, 2, size=(CHUNK, net.n)).astype(np.uint8)  # genRandState (base.py:368-370)
        ints = rng.integers(0, net.n, size=(CHUNK, ITERS)).astype(np.int32)  # Graph.step: random.randint(0, N-1)
        dbls = rng.random(size=(CHUNK, ITERS * (net.n + 1)))  # per iteration: N flip draws, then the predictor pick
        hist += orc.ssd(net, env, st, ITERS, P, TGT, orc.Draws(ints=ints, dbls=dbls)).astype(np.int64)
    return hist


if __name__ == "__main__":
    with mp.get_context("spawn").Pool(min(8, mp.cpu_count())) as pool:
        hists = pool.map(batch, range(BATCHES))
    h = np.array(hists)
    assert h.sum() == BATCHES * CHAINS * ITERS
    np.savez_compressed(ROOT / "tests" / "golden" / "b100_ssd_literal.npz", hist=h, iters=ITERS, chains=CHAINS, batches=BATCHES,
                        p=P, pickle=np.array(NAME), tgt_nodes=TGT)
    print("total iterations", h.sum(), "top buckets", np.argsort(-h.sum(0))[:5], (h.sum(0) / h.sum())[np.argsort(-h.sum(0))[:5]])
