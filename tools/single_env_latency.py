"""Wall-clock latency of the single-env drop-in API (one env.step = one launch + one read-back)."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "gym-pbn-stac_b200"))
import gym_PBN  # noqa: E402

EX5 = (["u", "x1", "x2", "x3", "x4"],
       [[], [("not x2 and not x4", 1)], [("not x4 and not u and (x2 or x3)", 1)],
        [("not x2 and not x4 and x1", 0.7), ("False", 0.3)], [("not x2 and not x3", 1)]])
GOAL = {"target_nodes": {(0, 0, 0, 0, 1)}, "target": {(0, 0, 0, 0, 1)},
        "all_attractors": [{(0, 0, 1, 0, 0)}, {(0, 0, 0, 0, 1)}]}


def bench(env, action, n=2000):
    env.reset(seed=1)
    for _ in range(50):
        o, r, te, tr, _ = env.step(action)
        if te or tr:
            env.reset()
    t0 = time.perf_counter()
    for _ in range(n):
        o, r, te, tr, _ = env.step(action)
        if te or tr:
            env.reset()
    return (time.perf_counter() - t0) / n * 1e6


print("PBN-v0 example network: %.1f us per env.step" % bench(gym_PBN.make("gym-PBN/PBN-v0", logic_func_data=EX5, goal_config=dict(GOAL)), 2))
star = [[("*",) * 100], [("*",) * 100]]
print("Bittner-100 (all-attracting): %.1f us per env.step" % bench(gym_PBN.make("gym-PBN/Bittner-100-v0", all_attractors=star), 0))
print("Bittner-100 (sampled attractors): %.1f us per env.step" % bench(gym_PBN.make("gym-PBN/Bittner-100-v0", seed=1), 3, n=500))
