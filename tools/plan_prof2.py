"""Per-pass timing and parked counts of a multi-pass env.step plan (config 2)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gym-pbn-stac_b200"))
from gym_PBN.b200 import abi, compiler, engine  # noqa: E402
from gym_PBN.b200 import attractors as att_tools  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
plan = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [32, 256, 0]
which = sys.argv[3] if len(sys.argv) > 3 else "c2"
torch.manual_seed(0)
if which == "c2":
    net = engine.Network(compiler.load_bittner("28_15_median"))
    atts = att_tools.exact_attractor_cubes(net)
    env = engine.EnvImage(net, abi.ENV_TARGET, attractors=atts, horizon=100, max_inner=4096)
    acts = torch.randint(0, 29, (B, 1), dtype=torch.int32, device="cuda")
else:  # config 4: Bittner-200 multi on its verified attractor
    import json
    net = engine.Network(compiler.load_bittner("200_5_kmeans"))
    atts = [[tuple(c) for c in a] for a in json.loads((ROOT / "tests" / "golden" / "b200_verified_attractors.json").read_text())]
    env = engine.EnvImage(net, abi.ENV_MULTI, attractors=atts, horizon=100, max_inner=4096, dedup=True)
    acts = torch.randint(0, net.n + 1, (B, 3), dtype=torch.int32, device="cuda")
sim = engine.Simulator(net, B, seed=1)
sim.env_reset(env)
for it in range(4):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(plan) + 1)]
    parked = []
    evs[0].record()
    for k, b in enumerate(plan):
        if k == 0:
            sim.env_step(env, acts, budget=b)
        else:
            sim.env_step_resume(env, budget=b)
        evs[k + 1].record()
    torch.cuda.synchronize()
    # parked counts are in the work lists' headers: re-run is not needed, read running after the fact per pass is lost; report final
    ts = [evs[k].elapsed_time(evs[k + 1]) * 1e3 for k in range(len(plan))]
    print(f"iter {it}: passes {plan} us {[round(t, 1) for t in ts]} total {sum(ts):.1f}  running left {int(sim.running.sum())}")
    sim.env_reset(env, mask=(sim.terminated | sim.truncated))
# parked counts per pass (separate run with syncs)
sim.env_step(env, acts, budget=plan[0]); print("parked after pass 1:", int(sim.running.sum()))
for b in plan[1:]:
    sim.env_step_resume(env, budget=b); print(f"parked after resume({b}):", int(sim.running.sum()))
