"""Sampled + verified attractors: trap-space percolation, symbolic closure check, exact terminal SCCs inside
a trap space — host parts on CPU (states sampled with the oracle), the whole route on the GPU against the exhaustive search."""
import numpy as np
import pytest

import oracle as orc


def _model(name):
    from gym_PBN.b200 import compiler
    from gym_PBN.b200.attractors import SuccessorModel

    spec = compiler.load_bittner(name)
    return spec, SuccessorModel(spec)


def _sampled_states(name, n_states, seed):
    sets, ids = orc.load_bittner(name)
    onet = orc.net_from_predictor_sets(sets, ids)
    st = orc.rand_state(onet, n_states, orc.Draws(seed=seed, epoch=0))
    orc.rollout(onet, st, 64 * st.shape[1], orc.Draws(seed=seed, epoch=1))
    return onet, np.unique(st, axis=0)


def _states(cubes):
    from gym_PBN.b200.attractors import expand_cube

    return {s for c in cubes for s in expand_cube(c, limit=1 << 17)}


def test_trap_spaces_of_bittner28_hold_its_two_attractors():
    from gym_PBN.b200.attractors import cubes_closed, terminal_sccs_in_cube, trap_space

    spec, model = _model("28_15_median")
    onet, ends = _sampled_states("28_15_median", 48, 3)
    spaces = {trap_space(model, s) for s in ends}
    assert all(cubes_closed(model, [t]) for t in spaces)
    found = {}
    for t in spaces:
        if sum(v == "*" for v in t) <= 16:
            for cubes in terminal_sccs_in_cube(model, t, 16):
                assert cubes_closed(model, cubes)
                found[frozenset(_states(cubes))] = cubes
    sizes = sorted(len(k) for k in found)
    assert sizes == [120, 49152]  # the two attractors the exhaustive device search finds
    assert sorted(len(c) for c in found.values()) == [2, 4]
    # an attractor is closed under the ORACLE's dynamics too: a long rollout from inside never leaves it
    small = min(found, key=len)
    st = np.array([list(s) for s in list(small)[:64]], np.uint8)
    orc.rollout(onet, st, 2000, orc.Draws(seed=9, epoch=0))
    assert all(tuple(int(v) for v in row) in small for row in st)
    # a cube set with a hole is rejected
    cubes4 = max(found.values(), key=len)
    assert not cubes_closed(model, cubes4[:-1])


def test_closure_check_on_a_truth_table_network():
    from gym_PBN.b200 import compiler
    from gym_PBN.b200.attractors import SuccessorModel, cubes_closed, terminal_sccs_in_cube, trap_space
    from gym_PBN.utils.converters import logic_funcs_to_PBN_data

    names = ["u", "x1", "x2", "x3", "x4"]
    funcs = [[("u", 1.0)], [("x2 and x3", 0.5), ("x1", 0.5)], [("x1 or u", 1.0)], [("not x4", 0.7), ("x3", 0.3)], [("x4", 1.0)]]
    spec = compiler.compile_pbn_data(logic_funcs_to_PBN_data(names, funcs))
    model = SuccessorModel(spec)
    # brute force: explicit successor sets of all 32 states (node 0 is never updated, common/pbn.py:90)
    succ = {}
    for s in range(32):
        bits = [(s >> (4 - i)) & 1 for i in range(5)]
        out = set()
        for i in range(1, 5):
            c0, c1 = model.can(i, bits)
            for b, ok in ((0, c0), (1, c1)):
                if ok:
                    t = list(bits)
                    t[i] = b
                    out.add(tuple(t))
        succ[tuple(bits)] = out
    for s in succ:
        t = trap_space(model, s)
        inside = {x for x in succ if all(c == "*" or c == v for c, v in zip(t, x))}
        assert all(succ[x] <= inside for x in inside) and cubes_closed(model, [t])  # closed ...
        reach, todo = {s}, [s]
        while todo:
            for y in succ[todo.pop()]:
                if y not in reach:
                    reach.add(y)
                    todo.append(y)
        assert reach <= inside  # ... and contains everything reachable from the state
        for cubes in terminal_sccs_in_cube(model, t):
            members = _states(cubes)
            assert all(succ[x] <= members for x in members)


def test_restricted_network_equals_the_cube_dynamics():
    from gym_PBN.b200.attractors import SuccessorModel, restricted_network, terminal_sccs_in_cube, trap_space

    spec, model = _model("28_15_median")
    _, ends = _sampled_states("28_15_median", 24, 5)
    t = min((trap_space(model, s) for s in ends), key=lambda c: sum(v == "*" for v in c))
    sub, free = restricted_network(spec, t)
    a = terminal_sccs_in_cube(SuccessorModel(sub), tuple(["*"] * sub.n))
    b = terminal_sccs_in_cube(model, t)
    assert sorted(sorted(tuple(c[v] for v in free) for c in cubes) for cubes in b) == sorted(sorted(map(tuple, cubes)) for cubes in a)


@pytest.mark.gpu
def test_verified_route_equals_exhaustive_search_on_bittner28():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gym_PBN.b200 import attractors as at, compiler, engine

    net = engine.Network(compiler.load_bittner("28_15_median"))
    exact = at.exact_attractor_cubes(net)
    got, info = at.verified_attractors(net, resets=256, seed=1)
    assert sorted(len(_states(c)) for c in got) == sorted(len(_states(c)) for c in exact) == [120, 49152]
    assert {frozenset(_states(c)) for c in got} == {frozenset(_states(c)) for c in exact}
    assert all("exact" in i["method"] for i in info)
    # forcing the device search on the restricted network gives the same sets
    got2, info2 = at.verified_attractors(net, resets=256, seed=1, max_free_host=4)
    assert {frozenset(_states(c)) for c in got2} == {frozenset(_states(c)) for c in exact}
    assert any("device" in i["method"] for i in info2)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["100_5_kmeans", "200_5_kmeans"])
def test_verified_attractors_beyond_32_nodes_are_closed_on_the_device(name):
    """Bittner-100 / -200: every returned cube set is closed — checked symbolically by the route itself and here dynamically:
    envs reset INTO the attractors stay attracting under thousands of GPU updates (step-until-attractor never moves)."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gym_PBN.b200 import abi, attractors as at, compiler, engine

    net = engine.Network(compiler.load_bittner(name))
    atts, info = at.verified_attractors(net, resets=128, seed=2)
    assert len(atts) >= 1 and all(i["states"] >= 1 for i in info)
    model = at.SuccessorModel(net.spec)
    assert all(at.cubes_closed(model, c) for c in atts)
    two = atts if len(atts) >= 2 else atts * 2
    env = engine.EnvImage(net, abi.ENV_TARGET, attractors=two, horizon=100, max_inner=64)
    sim = engine.Simulator(net, 2048, seed=3)
    sim.env_reset(env)
    sim.rollout(3000)  # free-running dynamics from inside the attractors
    acts = torch.zeros((2048, 1), dtype=torch.int32, device="cuda")
    sim.env_step(env, acts)
    assert int(sim.inner.max()) == 1  # still attracting: the loop stopped after its mandatory first update
