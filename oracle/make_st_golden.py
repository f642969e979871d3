"""TEST INFRASTRUCTURE — golden traces of the reference's self-triggering envs (runs only in the build container):
python oracle/make_st_golden.py -> tests/golden/ex5_self_triggering.npz

Per env.step of PBNSelfTriggeringEnv / PBCNSelfTriggeringEnv (self_triggering.py:56-93,146-197) on the example network:
start state, action, the draws in call order (per primitive step: randint -> node, uniform -> node value, uniform -> stop),
and the outputs (observation, discounted reward, terminated, interval)."""
import contextlib
import io
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "oracle"))
import ref_loader  # noqa: E402

ns = ref_loader.load()
from gym_PBN.envs import self_triggering as st  # noqa: E402  (the reference's, after ref_loader.load())

EX5 = (["u", "x1", "x2", "x3", "x4"],
       [[], [("not x2 and not x4", 1)], [("not x4 and not u and (x2 or x3)", 1)],
        [("not x2 and not x4 and x1", 0.7), ("False", 0.3)], [("not x2 and not x3", 1)]])
GOAL = {"target_nodes": {(0, 0, 0, 0, 1)}, "target": {(0, 0, 0, 0, 1)}, "all_attractors": [{(0, 0, 1, 0, 0)}, {(0, 0, 0, 0, 1)}]}
quiet = lambda: contextlib.redirect_stdout(io.StringIO())  # noqa: E731

out = {}
for tag, cls, kw in (("pbn", st.PBNSelfTriggeringEnv, dict(T=5)), ("pbcn", st.PBCNSelfTriggeringEnv, dict(T=7))):
    with quiet():
        env = cls(logic_func_data=EX5, goal_config=dict(GOAL), gamma=0.9, **kw)
    rng = np.random.default_rng(11)
    starts, acts, ints_l, dbls_l, obs_l, rew_l, term_l, intv_l = [], [], [], [], [], [], [], []
    with ref_loader.Recorder() as rec, quiet():
        env.reset(seed=3)
        rec.take()
        for _ in range(60):
            start = rng.integers(0, 2, 5).astype(bool)
            env.PBN.state = start.copy()
            if tag == "pbn":
                action = (int(rng.integers(0, 6)), int(rng.integers(1, 11)))
                acts.append(action)
            else:
                # the flat index form: the tuple form dies in np.isreal on NumPy >= 1.24 (self_triggering.py:152)
                action = int(rng.integers(0, 20))
                acts.append((action % 2, action // 2 + 1))
            obs, r, term, trunc, info = env.step(action)
            ints, dbls = rec.take()
            assert len(ints) == info["interval"] and len(dbls) == 2 * info["interval"]
            starts.append(start); ints_l.append(ints + [0] * (8 - len(ints))); dbls_l.append(dbls + [0.0] * (16 - len(dbls)))
            obs_l.append(np.array(obs).astype(np.uint8)); rew_l.append(float(r)); term_l.append(bool(term)); intv_l.append(info["interval"])
    out.update({f"{tag}_start": np.array(starts, np.uint8), f"{tag}_action": np.array(acts, np.int32),
                f"{tag}_ints": np.array(ints_l, np.int32), f"{tag}_dbls": np.array(dbls_l, np.float64),
                f"{tag}_obs": np.array(obs_l), f"{tag}_reward": np.array(rew_l), f"{tag}_term": np.array(term_l),
                f"{tag}_interval": np.array(intv_l, np.int32)})
np.savez_compressed(ROOT / "tests" / "golden" / "ex5_self_triggering.npz", **out)
print({k: v.shape for k, v in out.items()}, out["pbn_interval"][:10], out["pbn_reward"][:5])
