"""TEST INFRASTRUCTURE — records golden vectors from the UNMODIFIED reference (container-only).

    python oracle/make_golden.py [--long]

Every fixture under tests/golden/ is produced here by running the reference's own classes
(imported from /root/reference through oracle/ref_loader.py) while recording each random draw
the hot path consumes.  The oracle (oracle/pbn_oracle.c) and the CUDA path are
then replayed against the same draws and must reproduce states/rewards bit for bit.

Trace schema (one .npz per scenario; T ops for each of E independent trajectories, key prefix "e{k}/"):
  op[T]            0 = env.step / core step, 1 = reset
  act[T,K]         action(s) of the op (unused for reset)
  int_off[T+1], ints[]   recorded integer draws consumed by op t = ints[int_off[t]:int_off[t+1]]
  dbl_off[T+1], dbls[]   recorded float64 draws, same layout
  state[T,N]       network state AFTER the op;  obs[T,N] what env.step returned
  reward[T], term[T], trunc[T], target_att[T]
"""
import argparse
import contextlib
import io
import json
import pickle
import random
import sys
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import ref_loader  # noqa: E402

GOLD = HERE.parent / "tests" / "golden"
DATA = HERE.parent / "gym-pbn-stac_b200" / "gym_PBN" / "envs" / "bittner" / "data"

EX5 = (
    ["u", "x1", "x2", "x3", "x4"],
    [
        [],
        [("not x2 and not x4", 1)],
        [("not x4 and not u and (x2 or x3)", 1)],
        [("not x2 and not x4 and x1", 0.7), ("False", 0.3)],
        [("not x2 and not x3", 1)],
    ],
)  # the network of example.py:26-35


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def load_sets(name):
    sets = pickle.load(open(DATA / f"predictor_sets_{name}.pkl", "rb"))
    ids = json.load(open(DATA / "node_ids.json"))[name]["node_ids"]
    return sets, ids


class Trace:
    def __init__(self, n, k):
        self.n, self.k = n, k
        self.op, self.act, self.state, self.obs = [], [], [], []
        self.reward, self.term, self.trunc, self.tatt = [], [], [], []
        self.ints, self.dbls, self.int_off, self.dbl_off = [], [], [0], [0]

    def add(self, op, act, draws, state, obs=None, reward=0, term=False, trunc=False, tatt=-1):
        ints, dbls = draws
        a = np.full(self.k, -1, np.int32)
        act = np.atleast_1d(np.asarray(act, np.int32))
        a[: len(act)] = act
        self.op.append(op), self.act.append(a)
        self.ints += ints
        self.dbls += dbls
        self.int_off.append(len(self.ints)), self.dbl_off.append(len(self.dbls))
        self.state.append(np.asarray(state, np.uint8))
        self.obs.append(np.asarray(state if obs is None else obs, np.uint8))
        self.reward.append(int(reward)), self.term.append(bool(term)), self.trunc.append(bool(trunc))
        self.tatt.append(int(tatt))

    def dump(self, prefix):
        return {
            f"{prefix}op": np.array(self.op, np.int8), f"{prefix}act": np.array(self.act, np.int32).reshape(-1, self.k),
            f"{prefix}int_off": np.array(self.int_off, np.int64), f"{prefix}ints": np.array(self.ints, np.int32),
            f"{prefix}dbl_off": np.array(self.dbl_off, np.int64), f"{prefix}dbls": np.array(self.dbls, np.float64),
            f"{prefix}state": np.array(self.state, np.uint8).reshape(-1, self.n),
            f"{prefix}obs": np.array(self.obs, np.uint8).reshape(-1, self.n),
            f"{prefix}reward": np.array(self.reward, np.int32), f"{prefix}term": np.array(self.term, np.uint8),
            f"{prefix}trunc": np.array(self.trunc, np.uint8), f"{prefix}target_att": np.array(self.tatt, np.int32),
        }


def cubes_to_arrays(attractors, n):
    cubes, off = [], [0]
    for att in attractors:
        for c in att:
            cubes.append([2 if v == "*" else int(v) for v in c])
        off.append(len(cubes))
    return np.array(cubes, np.int8).reshape(-1, n), np.array(off, np.int32)


# ------------------------------------------------------------------------------------------ truth-table family
def pbn_data_arrays(pbn_data):
    n = len(pbn_data)
    masks = np.array([np.asarray(d[0], bool) for d in pbn_data])
    kmax = int(masks.sum(1).max())
    tables = np.zeros((n, 2**kmax))
    for i, d in enumerate(pbn_data):
        t = np.asarray(d[1], np.float64).reshape(-1)
        tables[i, : t.size] = t
    return masks, tables


def gen_ex5_pbnenv(ns, steps=3000, seeds=(0, 1)):
    out = {}
    for e, seed in enumerate(seeds):
        with quiet():
            env = ns.pbn_env.PBNEnv(
                logic_func_data=EX5,
                goal_config={"target_nodes": {(0, 0, 0, 0, 1)}, "target": {(0, 0, 0, 0, 1)},
                             "all_attractors": [{(0, 0, 1, 0, 0)}, {(0, 0, 0, 0, 1)}]},
            )
        if e == 0:
            pbn_data = ns.converters.logic_funcs_to_PBN_data(*EX5)
            masks, tables = pbn_data_arrays(pbn_data)
            atts = [list(a) for a in env.all_attractors]  # iteration order of each set = order random.choice indexes
            cubes, off = cubes_to_arrays(atts, 5)
            out.update(masks=masks, tables=tables, att_cubes=cubes, att_off=off,
                       targets=np.array(sorted(env.target_nodes), np.int8),
                       attractors_sorted=np.array(sorted(tuple(s) for a in env.all_attractors for s in a), np.int8))
        tr = Trace(5, 1)
        rng = np.random.default_rng(123 + seed)
        with ref_loader.Recorder() as rec, quiet():
            obs, info = env.reset(seed=seed)
            ints, dbls = rec.take()
            tr.add(1, 0, (ints[1:], dbls), obs.copy())  # ints[0] is the discarded first choice (pbn_env.py:199)
            for _ in range(steps):
                a = int(rng.integers(0, 5))
                obs, r, term, trunc, info = env.step(a)
                tr.add(0, a, rec.take(), obs.copy(), reward=r, term=term, trunc=trunc)
                assert info["observation_idx"] == int("".join(str(int(x)) for x in obs), 2)
                if term:
                    obs, info = env.reset()
                    ints, dbls = rec.take()
                    tr.add(1, 0, (ints[1:], dbls), obs.copy())
        out.update(tr.dump(f"e{e}/"))
    out["n_traj"] = len(seeds)
    np.savez_compressed(GOLD / "ex5_pbnenv.npz", **out)


def gen_ex5_pbcn_sampled(ns, steps=400, seed=5):
    """PBCNEnv + PBCNSampledDataEnv on the example network (control ignored by the dynamics, Q4)."""
    goal = {"target_nodes": {(0, 0, 0, 0, 1)}, "target": {(0, 0, 0, 0, 1)},
            "all_attractors": [{(0, 0, 1, 0, 0)}, {(0, 0, 0, 0, 1)}]}
    out = {}
    pbn_data = ns.converters.logic_funcs_to_PBN_data(*EX5)
    masks, tables = pbn_data_arrays(pbn_data)
    for e, cls in enumerate((ns.pbcn_env.PBCNEnv, ns.sampled_data.PBCNSampledDataEnv, ns.sampled_data.PBNSampledDataEnv)):
        with quiet():
            env = cls(logic_func_data=EX5, goal_config=dict(goal), **({"T": 8} if e else {}))
        atts = [list(a) for a in env.all_attractors]
        cubes, off = cubes_to_arrays(atts, 5)
        k = 1 if e == 0 else 2
        tr = Trace(5, k)
        rng = np.random.default_rng(seed + e)
        with ref_loader.Recorder() as rec, quiet():
            obs, info = env.reset(seed=seed + e)
            ints, dbls = rec.take()
            tr.add(1, 0, (ints[1:], dbls), obs.copy())
            for _ in range(steps):
                if e == 0:
                    a = int(rng.integers(0, 5))
                    obs, r, term, trunc, info = env.step(a)
                    act = [a]
                elif e == 1:
                    ctrl, interval = [bool(rng.integers(0, 2))], int(rng.integers(1, 9))
                    # the (array, int) tuple form dies in np.isreal under NumPy 2 (sampled_data.py:145); use the flat index
                    obs, r, term, trunc, info = env.step(int(ctrl[0]) + 2 * (interval - 1))
                    assert info["interval"] == interval
                    act = [interval, int(ctrl[0])]
                else:
                    a, interval = int(rng.integers(0, 6)), int(rng.integers(1, 9))
                    obs, r, term, trunc, info = env.step((a, interval))
                    act = [a, interval]
                tr.add(0, act, rec.take(), obs.copy(), reward=r, term=term, trunc=trunc)
                if term:
                    obs, info = env.reset()
                    ints, dbls = rec.take()
                    tr.add(1, 0, (ints[1:], dbls), obs.copy())
        out.update(tr.dump(f"e{e}/"))
        out[f"e{e}/att_cubes"], out[f"e{e}/att_off"] = cubes, off
        out[f"e{e}/M"] = getattr(env.PBN, "M", 0)
        out[f"e{e}/successful_reward"], out[f"e{e}/wrong_attractor_cost"] = env.successful_reward, env.wrong_attractor_cost
    out.update(masks=masks, tables=tables, targets=np.array([(0, 0, 0, 0, 1)], np.int8), n_traj=3,
               classes=np.array(["PBCNEnv", "PBCNSampledDataEnv", "PBNSampledDataEnv"]))
    np.savez_compressed(GOLD / "ex5_pbcn_sampled.npz", **out)


def gen_tt_core(ns, n=40, steps=400, seed=11):
    """PBN.step / flip / reset on a synthetic N=40, k<=4 truth-table network (two 32-bit words)."""
    rng = np.random.default_rng(seed)
    pbn_data = []
    for i in range(n):
        k = int(rng.integers(0, 5))
        mask = np.zeros(n, bool)
        mask[rng.choice(n, size=k, replace=False)] = True
        f1, f2 = rng.integers(0, 2, 2**k), rng.integers(0, 2, 2**k)
        c = float(rng.uniform(0.1, 0.9))
        table = (c * f1 + (1 - c) * f2).reshape([2] * k) if k else np.array(float(rng.integers(0, 2)))
        pbn_data.append((mask, table, i, f"G{i}", False))
    with quiet():
        pbn = ns.pbn.PBN(PBN_data=pbn_data)
    masks, tables = pbn_data_arrays(pbn_data)
    tr = Trace(n, 1)
    random.seed(seed), np.random.seed(seed)
    with ref_loader.Recorder() as rec, quiet():
        st = pbn.reset()  # np.random.rand(N) > 0.5 then state[0] = 0  (common/pbn.py:65,77)
        tr.add(1, 0, rec.take(), st.copy())
        for t in range(steps):
            a = int(rng.integers(0, n))
            if t % 3 == 0:
                pbn.flip(a)
            pbn.step()
            tr.add(0, a if t % 3 == 0 else -1, rec.take(), pbn.state.copy())
    out = tr.dump("e0/")
    out.update(masks=masks, tables=tables, n_traj=1)
    np.savez_compressed(GOLD / "tt40_core.npz", **out)


def gen_attractors(ns, seed=17):
    """PBNEnv.compute_attractors (exhaustive async STG + networkx attracting_components, pbn_env.py:238-255) on a few
    random truth-table networks — pins the exhaustive attractor search."""
    rng = np.random.default_rng(seed)
    out = {"n_nets": 0}
    for k, n in enumerate((5, 6, 7, 8, 9)):
        pbn_data = []
        for i in range(n):
            kin = int(rng.integers(0, 4))
            mask = np.zeros(n, bool)
            mask[rng.choice(n, size=kin, replace=False)] = True
            f1, f2 = rng.integers(0, 2, 2**kin), rng.integers(0, 2, 2**kin)
            c = float(rng.choice([0.0, 0.3, 1.0]))
            table = (c * f1 + (1 - c) * f2).reshape([2] * kin) if kin else np.array(float(rng.integers(0, 2)))
            pbn_data.append((mask, table, i, f"G{i}", False))
        with quiet():
            pbn = ns.pbn.PBN(PBN_data=pbn_data)
            env = ns.pbn_env.PBNEnv.__new__(ns.pbn_env.PBNEnv)
            env.PBN = pbn
            env.render_mode = "human"
            atts = env.compute_attractors()
        masks, tables = pbn_data_arrays(pbn_data)
        flat = sorted(sorted(a) for a in atts)
        out[f"n{k}/masks"], out[f"n{k}/tables"] = masks, tables
        out[f"n{k}/att_sizes"] = np.array([len(a) for a in flat], np.int32)
        out[f"n{k}/att_states"] = np.array([s for a in flat for s in a], np.int8).reshape(-1, n)
        out["n_nets"] = k + 1
    np.savez_compressed(GOLD / "tt_attractors.npz", **out)


# ------------------------------------------------------------------------------------------ Bittner graph
def gen_graph_core(ns, name, E, steps, sync_steps, seed):
    sets, ids = load_sets(name)
    n = len(ids)
    out = {"pickle": np.array(name), "n_traj": E}
    for e in range(E):
        g = ref_loader.build_graph(sets, ids)
        random.seed(seed + e), np.random.seed(seed + e)
        tr = Trace(n, 1)
        with ref_loader.Recorder() as rec:
            g.genRandState()
            tr.add(1, 0, rec.take(), g.getState())
            for t in range(steps):
                a = -1
                if t % 5 == 0:
                    a = (t * 7 + e) % n
                    g.flipNode(a)
                st = g.step()
                tr.add(0, a, rec.take(), st)
            for t in range(sync_steps):
                g.synch_step()
                tr.add(2, -1, rec.take(), g.getState())
        out.update(tr.dump(f"e{e}/"))
    np.savez_compressed(GOLD / f"b{name.split('_')[0]}_graph_core.npz", **out)


def fixture_attractors(ns, name, n_care=6, seed=99):
    """Attractor cube fixture by the reference's own sampling recipe (statistical_attractors,
    pbn_target.py:546-558: resets x forced single updates, most-visited first), applied to the projection
    on the first n_care nodes so the cubes are reachable; the rest of each cube is '*'."""
    sets, ids = load_sets(name)
    n = len(ids)
    g = ref_loader.build_graph(sets, ids)
    random.seed(seed)
    counts = {}
    for _ in range(20):
        g.genRandState()
        for _ in range(500):
            st = g.step()
            counts[st[:n_care]] = counts.get(st[:n_care], 0) + 1
    top = sorted(counts.items(), key=lambda kv: kv[1], reverse=True)[:5]
    star = ("*",) * (n - n_care)
    atts = [[tuple(p) + star] for p, _ in top[:3]]
    # one attractor with two cubes (second has a wildcard inside the cared prefix) — exercises multi-cube matching
    p3, p4 = top[3][0], top[4][0]
    atts.append([tuple(p3) + star, tuple(p4[:2]) + ("*",) + tuple(p4[3:]) + star])
    return atts


def bind_env(ns, cls, sets, ids, atts, horizon, cap):
    g = ref_loader.build_graph(sets, ids)
    goal = {"target_nodes": ids[:7], "target_node_values": ((0,) * 7,), "undesired_node_values": tuple(),
            "intervene_on": ids[:7], "horizon": horizon}
    with quiet():
        env = cls(g, goal, render_mode="human", name="fixture")
    env.all_attractors = atts
    if cls is ns.pbn_target.PBNTargetEnv:
        real = types.MethodType(ns.pbn_target.Bittner7.is_attracting_state, env)
    else:
        # the multi env tests membership in the wildcard-expanded set (pbn_target_multi.py:438-454,489-492);
        # cube matching accepts exactly the same states and does not need the exponential expansion
        real = types.MethodType(ns.pbn_target.Bittner7.is_attracting_state, env)
        env.attractor_count = len(atts)
        env.probabilities = [1 / len(atts)] * len(atts)
    calls = {"n": 0}

    def capped(state):  # inner-step cap (Q16): report "attracting" once `cap` updates were made in this env.step
        calls["n"] += 1
        return real(state) or calls["n"] >= cap

    env.is_attracting_state = capped
    env._calls = calls
    return env


def gen_target_env(ns, name="28_15_median", E=3, steps=250, horizon=20, cap=64, seed=21):
    sets, ids = load_sets(name)
    n = len(ids)
    atts = fixture_attractors(ns, name)
    cubes, off = cubes_to_arrays(atts, n)
    out = {"pickle": np.array(name), "n_traj": E, "att_cubes": cubes, "att_off": off, "horizon": horizon, "cap": cap}
    inner_total = 0
    for e in range(E):
        env = bind_env(ns, ns.pbn_target.PBNTargetEnv, sets, ids, atts, horizon, cap)
        rng = np.random.default_rng(seed + e)
        tr = Trace(n, 1)
        force = e == E - 1  # last trajectory: step(force=True) — exactly one update per env.step
        with ref_loader.Recorder() as rec, quiet():
            (st, tg), info = env.reset(seed=seed + e)
            ints, dbls = rec.take()
            tr.add(1, 0, (ints, dbls), st, tatt=atts.index(env.target))
            for _ in range(steps):
                a = int(rng.integers(0, n + 1))
                env._calls["n"] = 0
                obs, r, term, trunc, info = env.step(a, force=force)
                d = rec.take()
                inner_total += len(d[0])
                tr.add(0, [a], d, env.graph.getState(), obs=obs, reward=r, term=term, trunc=trunc)
                if term or trunc:
                    (st, tg), info = env.reset()
                    tr.add(1, 0, rec.take(), st, tatt=atts.index(env.target))
        out.update(tr.dump(f"e{e}/"))
        out[f"e{e}/force"] = int(force)
    print("target env micro-steps recorded:", inner_total)
    np.savez_compressed(GOLD / f"b{name.split('_')[0]}_target_env.npz", **out)


def gen_multi_env(ns, name="28_15_median", E=2, steps=200, horizon=25, cap=64, seed=31):
    import torch

    sets, ids = load_sets(name)
    n = len(ids)
    atts = fixture_attractors(ns, name)
    cubes, off = cubes_to_arrays(atts, n)
    out = {"pickle": np.array(name), "n_traj": E, "att_cubes": cubes, "att_off": off, "horizon": horizon, "cap": cap}
    for e in range(E):
        env = bind_env(ns, ns.pbn_target_multi.PBNTargetMultiEnv, sets, ids, atts, horizon, cap)
        rng = np.random.default_rng(seed + e)
        tr = Trace(n, 3)
        tensor_actions = e == 1  # e0: python list (duplicates cancel, all counted); e1: tensor (unique()'d)
        with ref_loader.Recorder() as rec, quiet():
            (st, tg), info = env.reset(seed=seed + e)
            tr.add(1, 0, rec.take(), st, tatt=len(atts) - 1)
            for _ in range(steps):
                a = [int(x) for x in rng.integers(0, n + 1, 3)]
                if rng.random() < 0.3:
                    a[2] = a[0]
                env._calls["n"] = 0
                obs, r, term, trunc, info = env.step(torch.tensor(a) if tensor_actions else list(a))
                tr.add(0, a, rec.take(), env.graph.getState(), obs=obs, reward=r, term=term, trunc=trunc)
                if term or trunc:
                    (st, tg), info = env.reset()
                    tr.add(1, 0, rec.take(), st, tatt=len(atts) - 1)
        out.update(tr.dump(f"e{e}/"))
        out[f"e{e}/dedup"] = int(tensor_actions)
    np.savez_compressed(GOLD / f"b{name.split('_')[0]}_multi_env.npz", **out)


# ------------------------------------------------------------------------------------------ SSD
TARGET_IDS = [234237, 324901, 759948, 25485, 266361, 108208, 130057]  # pbn_target.py:447


def ssd_env(ns, name):
    sets, ids = load_sets(name)
    n = len(ids)
    atts = [[("*",) * n], [("*",) * n]]  # all-attracting fixture: one update per env.step; reset = uniform random state
    g = ref_loader.build_graph(sets, ids)
    goal = {"target_nodes": TARGET_IDS, "target_node_values": ((0,) * 7,), "undesired_node_values": tuple(),
            "intervene_on": TARGET_IDS[:1], "horizon": 10**9}
    with quiet():
        env = ns.pbn_target.PBNTargetEnv(g, goal, render_mode="human", name="ssd-fixture")
    env.all_attractors = atts
    env.is_attracting_state = types.MethodType(ns.pbn_target.Bittner7.is_attracting_state, env)
    return env, ids


def gen_ssd_replay(ns, name="100_5_kmeans", chains=2, iters=150, seed=41):
    env, ids = ssd_env(ns, name)
    n = len(ids)
    out = {"pickle": np.array(name), "chains": chains, "iters": iters, "p": 0.01,
           "tgt_nodes": np.array([ids.index(t) for t in TARGET_IDS], np.int32)}
    random.seed(seed), np.random.seed(seed)
    hists, ints, dbls, init = [], [], [], []
    for c in range(chains):
        with ref_loader.Recorder() as rec, quiet():
            env.reset()  # _ssd_run resets itself (eval.py:78); do it here so the start state can be recorded
            rec.take()
            init.append(np.array(env.graph.getState(), np.uint8))
            orig_reset, env.reset = env.reset, (lambda *a, **k: None)
            h = ns.eval._ssd_run(7, iters, 0.01, None, env)
            env.reset = orig_reset
            i, d = rec.take()
        hists.append(h.astype(np.int64)), ints.append(i), dbls.append(d)
    out.update(hist=np.array(hists), init=np.array(init),
               ints=np.array(ints, np.int32), dbls=np.array(dbls, np.float64))
    np.savez_compressed(GOLD / "b100_ssd_replay.npz", **out)


class GoldenPolicy:
    """Deterministic stand-in for an agent: flip the first target gene that is 0 (1-indexed action), else no action."""

    def __init__(self, tgt_idx):
        self.tgt_idx = list(tgt_idx)

    def predict(self, state, target, deterministic=True):
        for i in self.tgt_idx:
            if int(state[i]) == 0:
                return (i + 1, None)
        return (0, None)


def gen_ssd_policy(ns, name="28_15_median", iters=250, horizon=10**9, cap=64, seed=71):
    """The `model` branch of _ssd_run (utils/eval.py:97-101) on Bittner-28 with the attractor fixture."""
    sets, ids = load_sets(name)
    n = len(ids)
    atts = fixture_attractors(ns, name)
    cubes, off = cubes_to_arrays(atts, n)
    env = bind_env(ns, ns.pbn_target.PBNTargetEnv, sets, ids, atts, horizon, cap)
    env.target_nodes = ids[:5]
    tgt_idx = list(range(5))
    random.seed(seed), np.random.seed(seed)
    int_off, dbl_off, ints, dbls, actions = [0], [0], [], [], []
    with ref_loader.Recorder() as rec, quiet():
        env.reset()
        rec.take()
        init = np.array(env.graph.getState(), np.uint8)
        tatt = atts.index(env.target)
        orig_step = env.step

        def step(action=0, force=False):
            env._calls["n"] = 0
            out = orig_step(action, force)
            i, d = rec.take()
            ints.extend(i), dbls.extend(d)
            int_off.append(len(ints)), dbl_off.append(len(dbls)), actions.append(int(action))
            return out

        env.step = step
        env.reset = lambda *a, **k: None
        h = ns.eval._ssd_run(5, iters, 0.01, GoldenPolicy(tgt_idx), env)
    np.savez_compressed(GOLD / "b28_ssd_policy.npz", pickle=np.array(name), att_cubes=cubes, att_off=off, cap=cap,
                        init=init, target_att=tatt, hist=h.astype(np.int64), tgt_nodes=np.array(tgt_idx, np.int32),
                        actions=np.array(actions, np.int32), ints=np.array(ints, np.int32), dbls=np.array(dbls, np.float64),
                        int_off=np.array(int_off, np.int64), dbl_off=np.array(dbl_off, np.int64),
                        final=np.array(env.graph.getState(), np.uint8))


def gen_ssd_long(ns, name="100_5_kmeans", iters=1_200_000, resets=300):
    """Two independent full-size reference estimates (utils/eval.py:20-72 defaults) — the TV yardstick."""
    import pandas as pd  # noqa: F401

    ns.eval.visualize_ssd = lambda *a, **k: None
    hs = []
    for run, seed in enumerate((1001, 2002)):
        env, ids = ssd_env(ns, name)
        random.seed(seed), np.random.seed(seed)
        with quiet():
            df, _ = ns.eval.compute_ssd_hist(env, iters=iters, resets=resets, bit_flip_prob=0.01, multiprocess=False)
        hs.append(df["Value"].to_numpy().astype(np.float64))
        print("ssd long run", run, "done; sum =", hs[-1].sum(), flush=True)
    np.savez_compressed(GOLD / "b100_ssd_long.npz", pickle=np.array(name), ssd=np.array(hs), iters=iters, resets=resets, p=0.01)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--long", action="store_true", help="also run the two full-size reference SSD estimates (~minutes)")
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    GOLD.mkdir(parents=True, exist_ok=True)
    ns = ref_loader.load()
    jobs = {
        "ex5": lambda: gen_ex5_pbnenv(ns),
        "pbcn": lambda: gen_ex5_pbcn_sampled(ns),
        "tt40": lambda: gen_tt_core(ns),
        "attractors": lambda: gen_attractors(ns),
        "g28": lambda: gen_graph_core(ns, "28_15_median", E=3, steps=400, sync_steps=20, seed=7),
        "g100": lambda: gen_graph_core(ns, "100_5_kmeans", E=2, steps=400, sync_steps=10, seed=8),
        "g200": lambda: gen_graph_core(ns, "200_5_kmeans", E=1, steps=300, sync_steps=5, seed=9),
        "target": lambda: gen_target_env(ns),
        "multi": lambda: gen_multi_env(ns),
        "target100": lambda: gen_target_env(ns, name="100_5_kmeans", E=2, steps=120, horizon=15, cap=48, seed=51),
        "multi100": lambda: gen_multi_env(ns, name="100_5_kmeans", E=2, steps=100, horizon=15, cap=48, seed=61),
        "ssd": lambda: gen_ssd_replay(ns),
        "ssd_policy": lambda: gen_ssd_policy(ns),
    }
    for k, fn in jobs.items():
        if args.only in (None, k):
            fn()
            print("wrote", k, flush=True)
    if args.long or args.only == "long":
        gen_ssd_long(ns)


if __name__ == "__main__":
    main()
