#!/usr/bin/env python
"""Times the UNMODIFIED Python reference (gym-PBN-stac, staged under baseline/_ref by __graft_entry__.build / tools/
stage_reference.py) on this box's host cores, next to the C port bench.py uses as the CPU arm.

    python baseline/time_reference.py [--seconds S] [--procs P]

One process per host core (multiprocessing spawn), each after a short warm-up runs for S/2 seconds
  * `Graph.step()` (gym_PBN/envs/bittner/base.py:306-312) on the shipped Bittner-100 predictor set, and
  * the body of `_ssd_run` (gym_PBN/utils/eval.py:76-103: env.render, getTargetIdx, np.random.rand(N) < p, flipNode,
    env.step(0)) on a PBNTargetEnv over the same graph with the all-attracting fixture,
and the last line printed is one JSON object with per-core and aggregate rates.  The reference needs two things to run at
all here, both applied by oracle/ref_loader.py without touching its sources: stub modules for packages the image lacks
(gymnasium, colomoto, matplotlib, plotly) and the one-method `Graph.getState` shim.  TEST INFRASTRUCTURE:
only bench.py's reference arm runs this."""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "baseline" / "_ref"
TARGET_IDS = [234237, 324901, 759948, 25485, 266361, 108208, 130057]  # pbn_target.py:447


def worker(seconds, q):
    import contextlib
    import io

    os.environ["GYM_PBN_REF"] = str(REF)
    sys.path.insert(0, str(ROOT / "oracle"))
    import numpy as np
    import oracle as orc  # only for the shipped pickle + node order (data loading)
    import ref_loader

    ns = ref_loader.load()
    sets, ids = orc.load_bittner("100_5_kmeans")
    g = ref_loader.build_graph(sets, ids)
    g.genRandState()
    t_end = time.perf_counter() + 0.5
    while time.perf_counter() < t_end:
        g.step()
    n, t0 = 0, time.perf_counter()
    t_end = t0 + seconds / 2
    while time.perf_counter() < t_end:
        for _ in range(200):
            g.step()
        n += 200
    step_rate = n / (time.perf_counter() - t0)
    # the _ssd_run loop body on the env the golden traces were recorded from (oracle/make_golden.py: ssd_env)
    nn = len(ids)
    goal = {"target_nodes": TARGET_IDS, "target_node_values": ((0,) * 7,), "undesired_node_values": tuple(),
            "intervene_on": TARGET_IDS[:1], "horizon": 10**9}
    with contextlib.redirect_stdout(io.StringIO()):
        env = ns.pbn_target.PBNTargetEnv(g, goal, render_mode="human", name="ssd-fixture")
    env.all_attractors = [[("*",) * nn], [("*",) * nn]]
    env.is_attracting_state = types.MethodType(ns.pbn_target.Bittner7.is_attracting_state, env)
    with contextlib.redirect_stdout(io.StringIO()):
        ns.eval._ssd_run(7, 50, 0.01, None, env)  # warm
        it, t0 = 0, time.perf_counter()
        t_end = t0 + seconds / 2
        while time.perf_counter() < t_end:
            ns.eval._ssd_run(7, 200, 0.01, None, env)
            it += 200
    ssd_rate = it / (time.perf_counter() - t0)
    q.put((step_rate, ssd_rate))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    args = ap.parse_args()
    if not (REF / "gym_PBN").is_dir():
        print(json.dumps({"error": "baseline/_ref is not staged"}))
        return
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(args.seconds, q)) for _ in range(args.procs)]
    for p in procs:
        p.start()
    res = [q.get(timeout=args.seconds * 4 + 100) for _ in procs]
    for p in procs:
        p.join()
    step = [r[0] for r in res]
    ssd = [r[1] for r in res]
    print(json.dumps({
        "kind": "reference", "what": "unmodified gym-PBN-stac Python reference, Bittner-100 (100_5_kmeans), one process per core",
        "cores": args.procs, "seconds_per_loop": args.seconds / 2,
        "graph_step_node_updates_per_s_per_core": sum(step) / len(step), "graph_step_node_updates_per_s": sum(step),
        "ssd_iterations_per_s_per_core": sum(ssd) / len(ssd), "ssd_iterations_per_s": sum(ssd),
        "ssd_1e10_iterations_days_on_this_box": 1.0e10 / sum(ssd) / 86400.0}))


if __name__ == "__main__":
    main()
