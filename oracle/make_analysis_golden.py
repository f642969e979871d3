"""TEST INFRASTRUCTURE — golden vectors of the reference's host-side analysis helpers on Bittner graphs (runs only in the
build container):  python oracle/make_analysis_golden.py  ->  tests/golden/graph_analysis.npz

  * 28-gene shipped set: Node.getStateProbs of every node and Graph.getNextStates at four states (base.py:66-87,221-242),
    and 60 forced-node updates Graph.step(i=...) under random.seed(5) (Node.Predstep, base.py:89-119);
  * a synthetic 6-gene predictor graph: Graph.genSTG + findAttractors (base.py:199-218,398-399) and sync_getNextStates.
"""
import contextlib
import io
import json
import pickle
import random
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "oracle"))
import ref_loader  # noqa: E402

DATA = ROOT / "gym-pbn-stac_b200" / "gym_PBN" / "envs" / "bittner" / "data"
ns = ref_loader.load()
out = {}

sets = pickle.load(open(DATA / "predictor_sets_28_15_median.pkl", "rb"))
ids = json.load(open(DATA / "node_ids.json"))["28_15_median"]["node_ids"]
g = ref_loader.build_graph(sets, ids)
rng = np.random.default_rng(3)
states = rng.integers(0, 2, size=(4, 28))
probs, nxt_states, nxt_probs, nxt_off = [], [], [], [0]
for s in states:
    g.setState(list(s))
    probs.append([node.getStateProbs(g.getState()) for node in g.nodes])
    d = g.getNextStates()
    for k in sorted(d):
        nxt_states.append(k)
        nxt_probs.append(d[k])
    nxt_off.append(len(nxt_states))
out.update(b28_states=states, b28_probs=np.array(probs), b28_next_states=np.array(nxt_states), b28_next_probs=np.array(nxt_probs),
           b28_next_off=np.array(nxt_off))
g.setState(list(states[0]))
random.seed(5)
trace = []
for k in range(60):
    with contextlib.redirect_stdout(io.StringIO()):
        trace.append(list(g.step(i=(7 * k) % 28)))
out["b28_forced_trace"] = np.array(trace)

# synthetic 6-gene graph
n, F = 6, 3
sid = [101, 205, 309, 412, 518, 623]
psets = []
cod = np.zeros((n, F)); A = np.zeros((n, F, 4)); inp = np.zeros((n, F, 3), np.int64)
for i in range(n):
    buf = np.empty((3, F), dtype=object)
    others = [x for x in range(n) if x != i]
    for f in range(F):
        trio = rng.choice(others, 3, replace=False)
        a = rng.normal(size=(4, 1)).round(2)
        c = float(rng.uniform(0.2, 1.0))
        buf[0, f], buf[1, f], buf[2, f] = c, a, np.array([sid[t] for t in trio])
        cod[i, f], A[i, f], inp[i, f] = c, a[:, 0], trio
    psets.append(buf)
g6 = ref_loader.build_graph(psets, sid)
with contextlib.redirect_stdout(io.StringIO()):
    stg = g6.genSTG()
atts = [sorted(a) for a in ns.base.findAttractors(stg)]
atts.sort()
out.update(s6_ids=np.array(sid), s6_cod=cod, s6_A=A, s6_inp=inp, s6_edges=np.array(sorted((u + v) for u, v in stg.edges())),
           s6_att_states=np.array([s for a in atts for s in a]), s6_att_off=np.cumsum([0] + [len(a) for a in atts]))
g6.setState([1, 0, 1, 1, 0, 0])
d = g6.sync_getNextStates()
out.update(s6_sync_states=np.array(sorted(d)), s6_sync_probs=np.array([d[k] for k in sorted(d)]))
np.savez_compressed(ROOT / "tests" / "golden" / "graph_analysis.npz", **out)
print({k: v.shape for k, v in out.items()})
