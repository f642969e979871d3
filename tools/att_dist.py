"""Distribution of inner updates per env.step for BASELINE config 2 (Bittner-28, exact attractors)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gym-pbn-stac_b200"))
from gym_PBN.b200 import abi, compiler, engine  # noqa: E402
from gym_PBN.b200 import attractors as att_tools  # noqa: E402

net = engine.Network(compiler.load_bittner("28_15_median"))
atts = att_tools.exact_attractor_cubes(net)
print("cubes per attractor:", [len(a) for a in atts])
env = engine.EnvImage(net, abi.ENV_TARGET, attractors=atts, horizon=100, max_inner=4096)
B = 65536
sim = engine.Simulator(net, B, seed=1)
sim.env_reset(env)
acts = torch.randint(0, 29, (B, 1), dtype=torch.int32, device="cuda")
for it in range(3):
    sim.env_step(env, acts)
    x = sim.inner.float()
    q = torch.quantile(x, torch.tensor([0.25, 0.5, 0.75, 0.9, 0.95, 0.96, 0.97, 0.99], device="cuda"))
    print(f"step {it}: mean {x.mean():.1f} quantiles 25/50/75/90/95/96/97/99 = {[int(v) for v in q]} capped {int((x >= 4096).sum())}")
    sim.env_reset(env, mask=(sim.terminated | sim.truncated))
