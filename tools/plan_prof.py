"""Per-pass timing of the two-pass env.step (budgeted first pass + resume pass) for BASELINE configs 2 and 4."""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gym-pbn-stac_b200"))
sys.path.insert(0, str(ROOT / "tools"))
from gym_PBN.b200 import abi, compiler, engine  # noqa: E402
from gym_PBN.b200 import attractors as att_tools  # noqa: E402
from bench_configs import cube_fixture  # noqa: E402


def prof_plans(label, net, env, B, acts, plans):
    for plan in plans:
        sim = engine.Simulator(net, B, seed=1)
        sim.plan_budgets = plan
        sim.env_reset(env)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t = 0.0
        reps = 10
        for it in range(reps + 2):
            a.record(); sim.env_step(env, acts); b.record()
            torch.cuda.synchronize()
            if it >= 2:
                t += a.elapsed_time(b)
            sim.env_reset(env, mask=(sim.terminated | sim.truncated))
        print(f"{label} B={B} plan={plan}: {t/reps*1e3:7.1f} us per step")


def prof(label, net, env, B, acts, budgets):
    for p1 in budgets:
        sim = engine.Simulator(net, B, seed=1)
        sim.env_reset(env)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        t1 = t2 = 0.0
        parked = 0
        reps = 8
        for it in range(reps + 2):
            if p1 == 0:
                sim.plan_budgets = ()
                ev[0].record(); sim.env_step(env, acts); ev[1].record(); ev[2].record()
            else:
                ev[0].record(); sim.env_step(env, acts, budget=p1); ev[1].record(); sim.env_step_resume(env); ev[2].record()
            torch.cuda.synchronize()
            if it >= 2:
                t1 += ev[0].elapsed_time(ev[1]); t2 += ev[1].elapsed_time(ev[2])
            sim.env_reset(env, mask=(sim.terminated | sim.truncated))
        print(f"{label} B={B} budget={p1:4d}: pass1 {t1/reps*1e3:7.1f} us  pass2 {t2/reps*1e3:7.1f} us  total {(t1+t2)/reps*1e3:7.1f} us  mean inner {float(sim.inner.float().mean()):.1f}")


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    budgets = (0, 32)
    plans = ((), (32,), (32, 256), (32, 128), (16, 128), (32, 512), (32, 128, 512), (16, 64, 256, 1024))
    if which in ("all", "c2"):
        net = engine.Network(compiler.load_bittner("28_15_median"))
        atts = att_tools.exact_attractor_cubes(net)
        env = engine.EnvImage(net, abi.ENV_TARGET, attractors=atts, horizon=100, max_inner=4096)
        for B in (65536, 1 << 20):
            torch.manual_seed(0)
            acts = torch.randint(0, 29, (B, 1), dtype=torch.int32, device="cuda")
            prof("config2", net, env, B, acts, budgets)
            prof_plans("config2", net, env, B, acts, plans)
    if which in ("all", "c4"):
        net = engine.Network(compiler.load_bittner("200_5_kmeans"))
        rng = np.random.default_rng(0)
        atts = cube_fixture(net.n, rng, care=4)
        env = engine.EnvImage(net, abi.ENV_MULTI, attractors=atts, horizon=100, max_inner=4096, dedup=True)
        B = 131072
        torch.manual_seed(0)
        acts = torch.randint(0, net.n + 1, (B, 3), dtype=torch.int32, device="cuda")
        prof("config4", net, env, B, acts, budgets)
        prof_plans("config4", net, env, B, acts, plans)
