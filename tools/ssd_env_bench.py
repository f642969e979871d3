"""SSD estimation with the step-until-attractor loop inside every iteration (env given): Bittner-28 with its exact
attractors, cap 4 096, random start states.  python tools/ssd_env_bench.py"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / 'gym-pbn-stac_b200'))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from gym_PBN.b200 import abi, attractors, compiler, engine  # noqa: E402

net = engine.Network(compiler.load_bittner("28_15_median"))
atts = attractors.exact_attractor_cubes(net)
env = engine.EnvImage(net, abi.ENV_TARGET, attractors=atts, horizon=100, max_inner=4096)
for B in (1 << 16, 1 << 20):
    sim = engine.Simulator(net, B, seed=1); sim.rand_state()
    tgt = np.arange(7, dtype=np.int32)
    sim.ssd(5, 0.01, tgt, env=env); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 20
    a.record(); sim.ssd(iters, 0.01, tgt, env=env); b.record(); torch.cuda.synchronize()
    t = a.elapsed_time(b) * 1e-3
    print(f"B={B}: SSD with attractor loop (Bittner-28 exact attractors, cap 4096): {B*iters/t:.4g} iterations/s, {t*1e3/iters:.2f} ms per iteration")
