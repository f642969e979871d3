"""CPU: host-side logic of the product — logic-expression compiler, PBN_data front end, predictor LUT tabulation,
cube compilation, registry, sharding — against known answers derived from the reference (SURVEY.md §4 KAT table)."""
import numpy as np
import pytest

import oracle as orc
from golden_util import load, pbn_data_from

EX5 = (["u", "x1", "x2", "x3", "x4"],
       [[], [("not x2 and not x4", 1)], [("not x4 and not u and (x2 or x3)", 1)],
        [("not x2 and not x4 and x1", 0.7), ("False", 0.3)], [("not x2 and not x3", 1)]])


def test_logic_evaluator_self_test():
    from gym_PBN.utils.logic.eval import LogicExpressionEvaluator

    ev = LogicExpressionEvaluator({"u": False, "x1": False, "x2": False, "x3": True, "x4": False})
    assert ev.evaluate("not x4 and not u and (x2 or x3)") is True  # utils/logic/eval.py:170-179
    assert LogicExpressionEvaluator.get_symbols("not x4 and not u and (x2 or x3)") == ["x4", "u", "x2", "x3"]
    assert ev.evaluate("x3 or x1 and x2") is True and ev.evaluate("(x3 or x1) and x2") is False
    assert ev.evaluate("not not x3") is True and ev.evaluate("False or not True") is False
    with pytest.raises(Exception):
        ev.evaluate("x3 and (x1")
    with pytest.raises(Exception):
        ev.evaluate("x9")
    with pytest.raises(Exception):
        ev.evaluate("")


def test_converter_matches_reference_tables():
    """Tables of the example network as the reference's converter produces them (recorded in the golden file)."""
    from gym_PBN.utils.converters import logic_funcs_to_PBN_data

    z = load("ex5_pbnenv.npz")
    ours = logic_funcs_to_PBN_data(*EX5)
    ref = pbn_data_from(z)
    for (m, t, name, ctrl), (rm, rt) in zip(ours, ref):
        assert np.array_equal(m, rm) and np.array_equal(t.reshape(-1), rt.reshape(-1))
    assert [c for *_x, c in ours] == [True, False, False, False, False]
    assert ours[3][1].reshape(-1).tolist() == [0, 0, 0, 0, 0.7, 0, 0, 0]


def test_compile_pbn_data_accepts_both_arities():
    from gym_PBN.b200 import compiler
    from gym_PBN.utils.converters import logic_funcs_to_PBN_data

    four = logic_funcs_to_PBN_data(*EX5)
    five = [(m, t, i, n, c) for i, (m, t, n, c) in enumerate(four)]
    a, b = compiler.compile_pbn_data(four), compiler.compile_pbn_data(five)
    for k in a.arrays:
        assert np.array_equal(a.arrays[k], b.arrays[k])
    assert a.first_updatable == 1 and a.control.tolist() == [True, False, False, False, False]
    with pytest.raises(ValueError):
        compiler.compile_pbn_data([(np.zeros(3, bool), np.zeros(4), "a", False)] * 3)


@pytest.mark.parametrize("name", ["28_15_median", "70_5_kmeans", "70_5_kmeans-log", "100_5_kmeans", "150_5_kmeans", "200_5_kmeans"])
def test_predictor_compiler_matches_oracle_tabulation(name):
    """The product compiler and the oracle tabulate LUTs / cumulative CODs independently; they must agree exactly."""
    from gym_PBN.b200 import compiler

    spec = compiler.load_bittner(name)
    sets, ids = orc.load_bittner(name)
    onet = orc.net_from_predictor_sets(sets, ids)
    for k in ("pr_off", "pr_in", "pr_lut", "pr_cum", "pr_codsum"):
        assert np.array_equal(spec.arrays[k], onet.a[k]), k
    assert spec.n == len(ids) and spec.first_updatable == 0
    if name == "100_5_kmeans":
        assert spec.arrays["pr_off"][-1] == 499 and spec.arrays["pr_off"][8] - spec.arrays["pr_off"][7] == 4


def test_predictor_lut_sign_convention():
    from gym_PBN.b200.compiler import predictor_lut16

    # Y = 0 iff X.A < 0: an all-zero input row gives X.A = 0 -> 1 (bittner/base.py:115-118)
    assert predictor_lut16(np.array([[1.0], [1.0], [1.0], [1.0]])) == 0xFFFF
    assert predictor_lut16(np.array([[-1.0], [-1.0], [-1.0], [-1.0]])) == 0x0001
    lut = predictor_lut16(np.array([[1.0], [0.0], [0.0], [-0.5]]))
    for idx in range(16):
        x0, x3 = (idx >> 3) & 1, idx & 1
        assert ((lut >> idx) & 1) == int(x0 - 0.5 * x3 >= 0)


def test_node_id_order_is_the_golden_pad_order():
    """IDs 0..69 of the 100-gene set = the golden list of the reference's own test (tests/test_bittner.py:27)."""
    from gym_PBN.b200 import compiler
    from gym_PBN.envs.bittner.utils import pad_ids
    import json

    golden70 = [234237, 324901, 759948, 25485, 266361, 108208, 130057, 357278, 39781, 49665, 39159, 23185, 417218, 31251,
                343072, 142076, 128100, 376725, 112500, 241530, 44563, 36950, 812276, 51018, 897806, 809473, 754538, 813533,
                161992, 306013, 418105, 841308, 53316, 427943, 45421, 471096, 44605, 471918, 280768, 510130, 470621, 38770,
                130100, 24588, 50043, 485690, 230360, 283617, 244086, 898092, 51740, 26789, 288733, 44584, 768272, 134829,
                51814, 363086, 364469, 770377, 110503, 193106, 25081, 767851, 244307, 254428, 142067, 25495, 526657, 50271]
    weights = json.load(open(compiler.DATA_DIR / "weighted_gene_ids.json"))
    assert len(weights) == 276 and len(set(weights)) == 244
    assert pad_ids(golden70[:7], 70, weights) == golden70
    assert compiler.load_bittner("100_5_kmeans").ids[:70] == golden70
    assert compiler.load_bittner("70_5_kmeans").ids == golden70


def test_compile_cubes():
    from gym_PBN.b200.compiler import compile_cubes

    cube, off, tf, nt = compile_cubes(4, [[(1, 0, "*", 1)], [(0, 0, 0, 0), ("*", "*", 1, 1)]], targets=[(1, 1, 1, 1)])
    assert cube.tolist() == [[1, 0, 2, 1], [0, 0, 0, 0], [2, 2, 1, 1], [1, 1, 1, 1]]
    assert off.tolist() == [0, 1, 3] and (tf, nt) == (3, 1)
    with pytest.raises(ValueError):
        compile_cubes(4, [[(1, 0, 1)]])


def test_parse_cabean_state_format():
    from gym_PBN.b200.attractors import cube_matches, expand_cube, parse_state

    assert parse_state("1 0 1 0 * * 1") == (1, 0, 1, 0, "*", "*", 1)
    # the report the reference keeps as its parser fixture (get_attractors_from_cabean.py:57-82)
    from gym_PBN.utils.get_attractors_from_cabean import parse_attractors, sample_cabean_out
    assert parse_state("1-0-1-0-----1-") == (1, 0, 1, 0, "*", "*", 1)
    assert parse_attractors(sample_cabean_out) == {0: [(1, 0, 1, 0, "*", "*", 1)], 1: [(1, 0, 1, 1, 1, 1, 0)],
                                                   2: [(1, 0, 1, 1, 1, 1, 1)], 3: [(1, 1, 1, 1, 1, 1, 0)]}
    assert len(expand_cube((1, "*", 0, "*"))) == 4 and (1, 1, 0, 0) in expand_cube((1, "*", 0, "*"))
    assert cube_matches((1, "*", 0), (1, 1, 0)) and not cube_matches((1, "*", 0), (0, 1, 0))


def test_registry_has_every_reference_id():
    import gym_PBN
    from gym_PBN.b200 import gym_compat

    ids = ["gym-PBN/PBN-v0", "gym-PBN/PBN-target-v0", "gym-PBN/PBCN-v0", "gym-PBN/PBN-sampled-data-v0",
           "gym-PBN/PBCN-sampled-data-v0", "gym-PBN/PBN-self-triggering-v0", "gym-PBN/PBCN-self-triggering-v0",
           "gym-PBN/PBN-target_multi-v0", "gym-PBN/BittnerMultiGeneral-v0"]
    ids += [f"gym-PBN/Bittner-{n}-v0" for n in (7, 28, 30, 70, 100, 200)]
    ids += [f"gym-PBN/BittnerMulti-{n}-v0" for n in (7, 10, 20, 25, 28, 30, 50)]
    if not gym_compat.HAVE_GYMNASIUM:
        for i in ids:
            assert i in gym_compat.registry, i
        assert gym_compat.registry["gym-PBN/Bittner-100-v0"]["max_episode_steps"] == 100


def test_spaces_stand_in():
    from gym_PBN.b200.gym_compat import spaces

    d = spaces.Discrete(5)
    assert d.contains(0) and d.contains(4) and not d.contains(5) and not d.contains([1])
    t = spaces.Tuple((spaces.Discrete(6), spaces.Discrete(8, start=1)))
    assert t.contains((5, 8)) and not t.contains((5, 0)) and not t.contains((6, 1))
    assert spaces.MultiBinary(3).contains([1, 0, 1]) and not spaces.MultiBinary(3).contains([1, 2, 1])


def test_shard_range_covers_everything():
    from gym_PBN.b200.dist import shard_range

    for total, world in ((10, 3), (1 << 20, 8), (5, 8), (0, 2)):
        spans = [shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    spans = [shard_range(1000, r, 3, align=32) for r in range(3)]
    assert spans == [(0, 352), (352, 704), (704, 1000)]


def test_product_never_imports_the_oracle():
    """The product path must not route through oracle/ (or any CPU fallback)."""
    import pathlib
    import re

    root = pathlib.Path(__file__).resolve().parent.parent / "gym-pbn-stac_b200"
    for f in list(root.rglob("*.py")) + list(root.rglob("*.cu")) + list(root.rglob("*.cuh")):
        text = f.read_text()
        assert not re.search(r"^\s*(import|from)\s+(oracle|ref_loader)\b", text, re.M), f
        assert "libpbn_oracle" not in text and "oracle/_build" not in text, f


def test_engine_fails_loudly_without_cuda():
    import torch

    from gym_PBN.b200 import abi, compiler, engine

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(abi.PbnError):
        engine.Network(compiler.load_bittner("28_15_median"))


def test_exhaustive_attractors_match_reference():
    """PBN.attractors (vectorised terminal-SCC search) == the reference's compute_attractors on recorded networks.
    Runs on the CPU: the attractor search is host code (constructor-time, not the step path)."""
    import types

    from gym_PBN.envs.common.node import Node
    from gym_PBN.envs.common.pbn import PBN

    z = load("tt_attractors.npz")
    for k in range(int(z["n_nets"])):
        masks, tables = z[f"n{k}/masks"], z[f"n{k}/tables"]
        n = len(masks)
        pbn = PBN.__new__(PBN)  # no device needed for the host-side search
        pbn.N = n
        pbn.nodes = np.empty(n, dtype=object)
        for i in range(n):
            kin = int(masks[i].sum())
            pbn.nodes[i] = Node(masks[i], tables[i, : 2**kin], i)
        got = sorted(sorted(a) for a in pbn.attractors_host())
        sizes = z[f"n{k}/att_sizes"]
        states = [tuple(int(v) for v in s) for s in z[f"n{k}/att_states"]]
        want, pos = [], 0
        for sz in sizes:
            want.append(states[pos:pos + sz])
            pos += sz
        assert got == want, k


def test_logic_compiler_agrees_with_python_semantics():
    """Random expressions over and/or/not/parentheses/True/False: the compiled closure tree == Python's own evaluation
    (same precedence: not > and > or), for every assignment of the symbols."""
    import itertools
    import random as pyrandom

    from gym_PBN.utils.logic.eval import LogicExpressionEvaluator, compile_expression

    rnd = pyrandom.Random(7)
    names = ["x1", "x2", "u", "geneA", "y10"]

    def gen(depth):
        if depth == 0 or rnd.random() < 0.25:
            return rnd.choice(names + ["True", "False"])
        kind = rnd.random()
        if kind < 0.25:
            return "not " + gen(depth - 1)
        if kind < 0.45:
            return "(" + gen(depth - 1) + ")"
        return gen(depth - 1) + rnd.choice([" and ", " or "]) + gen(depth - 1)

    for _ in range(300):
        expr = gen(4)
        fn = compile_expression(expr)
        syms = sorted(set(LogicExpressionEvaluator.get_symbols(expr)))
        for vals in itertools.product([False, True], repeat=len(syms)):
            env = dict(zip(syms, vals))
            assert bool(fn(env)) == bool(eval(expr, {"__builtins__": {}}, dict(env))), expr


def test_rollout_return_and_gae_math():
    """gym_PBN.b200.rollout: discounted returns and GAE on torch tensors against plain Python loops (CPU tensors)."""
    torch = pytest.importorskip("torch")
    from gym_PBN.b200.rollout import discounted_returns, gae_advantages

    g = torch.Generator().manual_seed(0)
    T, B, gamma, lam = 9, 5, 0.9, 0.8
    r = torch.randint(-5, 21, (T, B), generator=g).float()
    term = torch.rand(T, B, generator=g) < 0.2
    trunc = (torch.rand(T, B, generator=g) < 0.1) & ~term
    v = torch.randn(T + 1, B, generator=g)
    ret = discounted_returns(r, term | trunc, gamma, v[T])
    adv, target = gae_advantages(r, v, term, trunc, gamma, lam)
    for b in range(B):
        run, a = float(v[T, b]), 0.0
        for t in range(T - 1, -1, -1):
            done = bool(term[t, b] or trunc[t, b])
            run = float(r[t, b]) + gamma * run * (not done)
            assert abs(run - float(ret[t, b])) < 1e-4
            delta = float(r[t, b]) + gamma * float(v[t + 1, b]) * (not done) - float(v[t, b])
            a = delta + gamma * lam * (not done) * a
            assert abs(a - float(adv[t, b])) < 1e-4 and abs(a + float(v[t, b]) - float(target[t, b])) < 1e-4


def test_node_compute_next_value_draws_from_numpy_global():
    """common/node.py:2,37 — `from numpy import random`: the host-side draw is numpy.random.uniform, so a seeded replay of
    the reference's stream gives the reference's decisions."""
    from gym_PBN.envs.common.node import Node

    node = Node([True, True, False], np.array([[0.1, 0.9], [0.5, 0.3]]), 2, "c")
    st = np.array([1, 0, 1], bool)
    np.random.seed(1234)
    got = [bool(node.compute_next_value(st)) for _ in range(64)]
    np.random.seed(1234)
    want = [bool(np.random.uniform(0, 1) < 0.5) for _ in range(64)]
    assert got == want and node.get_next_value_prob(st) == 0.5
