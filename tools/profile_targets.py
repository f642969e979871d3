"""Short workload that launches the non-headline kernels once each (for ncu captures): async + sync rollout on Bittner-100,
the persistent step-until-attractor kernel on Bittner-28, the fused vector step."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gym-pbn-stac_b200"))
from gym_PBN.b200 import abi, compiler, engine  # noqa: E402

torch.manual_seed(0)
rng = np.random.default_rng(0)
net = engine.Network(compiler.load_bittner("100_5_kmeans"))
sim = engine.Simulator(net, 1 << 20, seed=1)
sim.rand_state()
sim.rollout(500)             # k_rollout async: 5.2e8 updates
sim.rollout(5, sync=True)    # k_rollout sync:  5.2e8 node updates
net28 = engine.Network(compiler.load_bittner("28_15_median"))
atts = []
for _ in range(4):
    c = ["*"] * 28
    for i in rng.choice(28, size=6, replace=False):
        c[i] = int(rng.integers(0, 2))
    atts.append([tuple(c)])
env = engine.EnvImage(net28, abi.ENV_TARGET, attractors=atts, horizon=100, max_inner=4096)
s28 = engine.Simulator(net28, 1 << 18, seed=1)
s28.env_reset(env)
acts = torch.randint(0, 29, (1 << 18, 1), dtype=torch.int32, device="cuda")
s28.env_step(env, acts)      # k_env_step_att
torch.cuda.synchronize()
print("inner mean", float(s28.inner.float().mean()), "updates", int(s28.inner.sum()))
