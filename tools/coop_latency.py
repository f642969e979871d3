"""Latency of one update in the step-until-attractor kernel: B envs that never reach an attractor (a cube no state
matches... here: one full-care cube) run to the cap; one warp, so time / cap = cycles per update at that group size."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gym-pbn-stac_b200"))
from gym_PBN.b200 import abi, compiler, engine  # noqa: E402


def run(name, Bs=(1, 2, 4, 8, 9, 32, 256), cap=4096, n_cubes=6):
    net = engine.Network(compiler.load_bittner(name))
    n = net.n
    # cubes nobody reaches in practice: full-care random states
    g = torch.Generator().manual_seed(0)
    atts = [[tuple(int(v) for v in torch.randint(0, 2, (n,), generator=g))] for _ in range(n_cubes)]
    env = engine.EnvImage(net, abi.ENV_TARGET, attractors=atts, horizon=100, max_inner=cap)
    for B in Bs:
        sim = engine.Simulator(net, B, seed=1)
        sim.rand_state()
        acts = torch.zeros((B, 1), dtype=torch.int32, device="cuda")
        for _ in range(3):
            sim.env_step(env, acts)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        reps = 10
        for _ in range(reps):
            sim.env_step(env, acts)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        inner = float(sim.inner.float().mean())
        print(f"{name} B={B:4d} cap={cap} mean_inner={inner:.0f}  {ms*1e3:.1f} us/launch  {ms*1e6/inner:.1f} ns/update = {ms*1e6/inner*1.965:.0f} cycles")


if __name__ == "__main__":
    run("28_15_median")
    run("100_5_kmeans", n_cubes=4)
    run("200_5_kmeans", n_cubes=4)
