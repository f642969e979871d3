"""TEST INFRASTRUCTURE — ctypes front end of oracle/pbn_oracle.c (the CPU restatement of the
reference's hot path).  Not product code: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg import this.  The product never falls back to it.

Networks are handed to the oracle in *reference form* (input masks + float64 tables, or the
predictor-set pickles' (COD, A, input IDs) triples); the oracle derives what it needs itself, with
the reference's own float expression for the predictor LUT (bittner/base.py:100-118) — it shares
no compiler code with the product.
"""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
BUILD = HERE / "_build"
LIB = BUILD / "libpbn_oracle.so"

ORC_TT, ORC_PRED = 0, 1
PHILOX, REPLAY = 0, 1
ENV_PBN, ENV_PBCN, ENV_TARGET, ENV_MULTI, ENV_PBN_SD, ENV_PBCN_SD, ENV_PBN_ST, ENV_PBCN_ST = range(8)


def build(force=False):
    src = HERE / "pbn_oracle.c"
    if force or not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        BUILD.mkdir(exist_ok=True)
        subprocess.check_call(
            ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-o", str(LIB), str(src), "-lm"]
        )
    return LIB


class OrcNet(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n", C.c_int32), ("first", C.c_int32),
                ("tt_in_off", C.c_void_p), ("tt_in", C.c_void_p), ("tt_tab_off", C.c_void_p), ("tt_prob", C.c_void_p),
                ("pr_off", C.c_void_p), ("pr_in", C.c_void_p), ("pr_lut", C.c_void_p),
                ("pr_cum", C.c_void_p), ("pr_codsum", C.c_void_p),
                ("tt_thr", C.c_void_p), ("pr_thr", C.c_void_p)]


class OrcDraws(C.Structure):
    _fields_ = [("mode", C.c_int32), ("epoch", C.c_uint32), ("seed", C.c_uint64),
                ("ints", C.c_void_p), ("dbls", C.c_void_p), ("int_stride", C.c_int64), ("dbl_stride", C.c_int64),
                ("used", C.c_void_p)]


class OrcEnv(C.Structure):
    _fields_ = [("kind", C.c_int32), ("horizon", C.c_int32), ("max_inner", C.c_int32), ("force", C.c_int32),
                ("dedup", C.c_int32), ("control_write", C.c_int32), ("n_control", C.c_int32),
                ("successful_reward", C.c_int32), ("wrong_attractor_cost", C.c_int32),
                ("n_att", C.c_int32), ("att_off", C.c_void_p), ("cube", C.c_void_p),
                ("tgt_first", C.c_int32), ("n_tgt", C.c_int32),
                ("gamma_pow", C.c_void_p), ("n_gamma", C.c_int32), ("max_interval", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
        _lib.orc_geom.restype = C.c_uint32
        _lib.orc_geom.argtypes = [C.c_uint32, C.c_float]
        _lib.orc_geom_inv.restype = C.c_float
        _lib.orc_geom_inv.argtypes = [C.c_double]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Net:
    """Reference-form network held as numpy arrays + the C struct pointing at them."""

    def __init__(self, kind, n, first, **arrs):
        self.kind, self.n, self.first = kind, n, first
        self.a = {k: np.ascontiguousarray(v) for k, v in arrs.items()}
        s = OrcNet(kind=kind, n=n, first=first)
        for k, v in self.a.items():
            setattr(s, k, _p(v))
        if kind == ORC_TT:
            self.a["tt_thr"] = np.zeros(len(self.a["tt_prob"]), np.uint32)
            s.tt_thr = _p(self.a["tt_thr"])
        else:
            self.a["pr_thr"] = np.zeros(len(self.a["pr_cum"]), np.uint32)
            s.pr_thr = _p(self.a["pr_thr"])
        self.c = s
        lib().orc_fill_thresholds(C.byref(s))


def net_from_pbn_data(pbn_data):
    """PBN_data tuples (input_mask, truth_table, ...): common/pbn.py:37-46, common/node.py:6-32.
    Inputs = masked node indices ascending; table flattened C-order (first masked node = MSB)."""
    n = len(pbn_data)
    in_off, ins, tab_off, prob = [0], [], [0], []
    for node in pbn_data:
        mask = np.asarray(node[0], dtype=bool)
        tab = np.asarray(node[1], dtype=np.float64)
        idx = np.nonzero(mask)[0]
        assert tab.size == 2 ** len(idx)
        ins += idx.tolist()
        in_off.append(len(ins))
        prob += tab.reshape(-1).tolist()
        tab_off.append(len(prob))
    return Net(ORC_TT, n, 1, tt_in_off=np.array(in_off, np.int32), tt_in=np.array(ins + [0], np.int32),
               tt_tab_off=np.array(tab_off, np.int32), tt_prob=np.array(prob, np.float64))


def predictor_lut(A):
    """bit (x0<<3|x1<<2|x2<<1|x3) = 0 if np.matmul(X.T, A) < 0. else 1, with X built as base.py:100-104 does."""
    lut = 0
    for idx in range(16):
        X = np.ones((4, 1))
        for j in range(4):
            X[j] = (idx >> (3 - j)) & 1
        y = np.matmul(X.T, A)
        if not (y < 0.0):
            lut |= 1 << idx
    return lut


def net_from_predictor_sets(sets, node_ids):
    """Predictor-set pickle (list of (3,F) object arrays: COD, A(4,1), input IDs) -> oracle net.
    Cumulative COD exactly as Node.add_predictors accumulates it (base.py:30-45)."""
    n = len(node_ids)
    pos = {int(g): i for i, g in enumerate(node_ids)}
    off, ins, luts, cums, sums = [0], [], [], [], []
    for i in range(n):
        codsum, prev, first = 0, None, True
        for cod, A, inp in sets[i].T:
            if cod is None:
                continue
            codsum += cod
            cur = cod if first else prev + cod
            first, prev = False, cur
            ins += [pos[int(g)] for g in inp] + [i]
            luts.append(predictor_lut(A))
            cums.append(cur)
        sums.append(codsum)
        off.append(len(luts))
    return Net(ORC_PRED, n, 0, pr_off=np.array(off, np.int32), pr_in=np.array(ins, np.int32),
               pr_lut=np.array(luts, np.uint16), pr_cum=np.array(cums, np.float64), pr_codsum=np.array(sums, np.float64))


class Env:
    def __init__(self, kind, n, attractors=(), targets=(), horizon=100, max_inner=1 << 30, force=0, dedup=1,
                 control_write=0, n_control=0, successful_reward=10, wrong_attractor_cost=2, gamma=None, max_interval=None):
        """attractors: list of lists of cubes (tuples over 0/1/'*'); targets: list of full states (PBN family);
        gamma / max_interval: self-triggering envs (gamma**i tabulated here with Python's float pow, as the reference does)."""
        cubes, off = [], [0]
        for att in attractors:
            for cube in att:
                cubes.append([2 if v == "*" else int(v) for v in cube])
            off.append(len(cubes))
        tgt_first = len(cubes)
        for t in targets:
            cubes.append([int(v) for v in t])
        self.cube = np.array(cubes, np.int8).reshape(-1, n) if cubes else np.zeros((1, n), np.int8)
        self.att_off = np.array(off, np.int32)
        self.c = OrcEnv(kind=kind, horizon=horizon, max_inner=max_inner, force=force, dedup=dedup,
                        control_write=control_write, n_control=n_control, successful_reward=successful_reward,
                        wrong_attractor_cost=wrong_attractor_cost, n_att=len(off) - 1, att_off=_p(self.att_off),
                        cube=_p(self.cube), tgt_first=tgt_first, n_tgt=len(targets))
        if gamma is not None:
            n_gamma = int(max_interval) if max_interval else 2048
            self.gamma_pow = np.array([float(gamma) ** i for i in range(n_gamma)], np.float64)
            self.c.gamma_pow, self.c.n_gamma, self.c.max_interval = _p(self.gamma_pow), n_gamma, int(max_interval or 0)


class Draws:
    def __init__(self, seed=None, epoch=0, ints=None, dbls=None, B=None):
        if ints is not None or dbls is not None:
            self.ints = np.ascontiguousarray(ints, np.int32) if ints is not None else np.zeros((B or 1, 1), np.int32)
            self.dbls = np.ascontiguousarray(dbls, np.float64) if dbls is not None else np.zeros((B or 1, 1), np.float64)
            nb = self.ints.shape[0]
            self.used = np.zeros((nb, 2), np.int64)
            self.c = OrcDraws(mode=REPLAY, ints=_p(self.ints), dbls=_p(self.dbls),
                              int_stride=self.ints.shape[1], dbl_stride=self.dbls.shape[1], used=_p(self.used))
        else:
            self.used = np.zeros((B, 2), np.int64) if B else None
            self.c = OrcDraws(mode=PHILOX, seed=int(seed), epoch=int(epoch), used=_p(self.used))


def philox(ctr, key):
    out = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
    return list(out)


def rollout(net, state, steps, draws, sync=False, env0=0):
    state = np.ascontiguousarray(state, np.uint8)
    lib().orc_rollout(C.byref(net.c), _p(state), C.c_int64(state.shape[0]), C.c_int64(env0), C.c_int64(steps),
                      C.c_int(int(sync)), C.byref(draws.c))
    return state


def rollout_sync_sliced(net, state, steps, draws, env0=0):
    state = np.ascontiguousarray(state, np.uint8)
    rc = lib().orc_rollout_sync_sliced(C.byref(net.c), _p(state), C.c_int64(state.shape[0]), C.c_int64(env0), C.c_int64(steps),
                                       C.byref(draws.c))
    if rc:
        raise ValueError(f"orc_rollout_sync_sliced: unsupported input (code {rc})")
    return state


def env_step(net, env, state, n_steps, target_att, actions, draws, env0=0):
    B = state.shape[0]
    actions = np.ascontiguousarray(actions, np.int32).reshape(B, -1)
    obs = np.zeros_like(state)
    reward = np.zeros(B, np.int32)
    term = np.zeros(B, np.uint8)
    trunc = np.zeros(B, np.uint8)
    inner = np.zeros(B, np.int32)
    lib().orc_env_step(C.byref(net.c), C.byref(env.c), _p(state), _p(n_steps), _p(target_att), _p(actions),
                       C.c_int(actions.shape[1]), _p(obs), _p(reward), _p(term), _p(trunc), _p(inner),
                       C.c_int64(B), C.c_int64(env0), C.byref(draws.c))
    return obs, reward, term, trunc, inner


def env_step_f64(net, env, state, actions, draws, env0=0):
    """Self-triggering envs: -> (obs, discounted reward float64, terminated, interval)."""
    B = state.shape[0]
    actions = np.ascontiguousarray(actions, np.int32).reshape(B, -1)
    obs = np.zeros_like(state)
    reward, rf = np.zeros(B, np.int32), np.zeros(B, np.float64)
    term, trunc, inner = np.zeros(B, np.uint8), np.zeros(B, np.uint8), np.zeros(B, np.int32)
    ns, ta = np.zeros(B, np.int32), np.zeros(B, np.int32)
    lib().orc_env_step_f64(C.byref(net.c), C.byref(env.c), _p(state), _p(ns), _p(ta), _p(actions), C.c_int(actions.shape[1]),
                           _p(obs), _p(reward), _p(rf), _p(term), _p(trunc), _p(inner), C.c_int64(B), C.c_int64(env0),
                           C.byref(draws.c))
    return obs, rf, term, inner


def env_reset(net, env, state, n_steps, target_att, draws, mask=None, target_state=None, env0=0):
    B = state.shape[0]
    if mask is not None:
        mask = np.ascontiguousarray(mask, np.uint8)
    lib().orc_env_reset(C.byref(net.c), C.byref(env.c), _p(state), _p(n_steps), _p(target_att), _p(target_state),
                        _p(mask), C.c_int64(B), C.c_int64(env0), C.byref(draws.c))


def sample_pair(prob, u):
    """np.random.choice(range(A), size=2, replace=False, p=prob) from the uniforms u (pbn_target_multi.py:232-235)."""
    prob = np.ascontiguousarray(prob, np.float64)
    u = np.ascontiguousarray(u, np.float64)
    ids = np.zeros(2, np.int32)
    lib().orc_sample_pair.restype = C.c_int
    used = lib().orc_sample_pair(_p(prob), C.c_int(len(prob)), _p(u), _p(ids))
    return (int(ids[0]), int(ids[1])), int(used)


def rework_probas(prob_row, s, t, episode_len):
    """In place on one float64 probability row (pbn_target_multi.py:159-181)."""
    lib().orc_rework_probas.restype = None
    lib().orc_rework_probas(_p(prob_row), C.c_int(len(prob_row)), C.c_int(int(s)), C.c_int(int(t)), C.c_int(int(episode_len)))


def env_reset_cur(net, env, state, n_steps, target_att, prob, pair_ids, draws, sample_pair=False, mask=None, target_state=None, env0=0):
    B = state.shape[0]
    if mask is not None:
        mask = np.ascontiguousarray(mask, np.uint8)
    rc = lib().orc_env_reset_cur(C.byref(net.c), C.byref(env.c), _p(state), _p(n_steps), _p(target_att), _p(target_state), _p(mask),
                                 _p(prob), _p(pair_ids), C.c_int(int(sample_pair)), C.c_int64(B), C.c_int64(env0), C.byref(draws.c))
    if rc:
        raise ValueError("orc_env_reset_cur: MULTI envs with 2..64 attractors, Philox draws")


def rand_state(net, B, draws, env0=0):
    state = np.zeros((B, net.n), np.uint8)
    lib().orc_rand_state(C.byref(net.c), _p(state), C.c_int64(B), C.c_int64(env0), C.byref(draws.c))
    return state


def ssd(net, env, state, iters, p, tgt_nodes, draws, env0=0):
    tgt = np.ascontiguousarray(tgt_nodes, np.int32)
    hist = np.zeros(1 << len(tgt), np.uint64)
    rc = lib().orc_ssd(C.byref(net.c), C.byref(env.c) if env is not None else None, _p(state), C.c_int64(state.shape[0]),
                       C.c_int64(env0), C.c_int64(iters), C.c_double(p), _p(tgt), C.c_int(len(tgt)), _p(hist),
                       C.byref(draws.c))
    if rc:
        raise ValueError("orc_ssd: env0 must be a multiple of 32 in Philox mode (groups of 32 chains share the flip stream)")
    return hist


def geom(r, p):
    return lib().orc_geom(C.c_uint32(int(r)), lib().orc_geom_inv(C.c_double(p)))


def num_threads():
    return lib().orc_num_threads()


# ------------------------------------------------------------------ shipped data (fixtures, not code)
DATA = HERE.parent / "gym-pbn-stac_b200" / "gym_PBN" / "envs" / "bittner" / "data"


def load_bittner(name):
    """(predictor sets, node ids) of a shipped pickle, e.g. name='100_5_kmeans'."""
    import json
    import pickle

    sets = pickle.load(open(DATA / f"predictor_sets_{name}.pkl", "rb"))
    ids = json.load(open(DATA / "node_ids.json"))[name]["node_ids"]
    return sets, ids


def geom_law(p, kmax=4096):
    """counts[k] = number of the 2^23 equally likely inputs that orc_geom maps to gap k (last bin = tail)."""
    counts = np.zeros(kmax, np.int64)
    lib().orc_geom_law(C.c_double(p), _p(counts), C.c_int(kmax))
    return counts
