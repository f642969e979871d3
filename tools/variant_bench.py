"""Times the SSD / rollout kernels of whichever library PBN_B200_LIB points at (kernel experiments)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gym-pbn-stac_b200"))
from gym_PBN.b200 import abi, compiler, engine  # noqa: E402


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    return best


name = sys.argv[1] if len(sys.argv) > 1 else "100_5_kmeans"
net = engine.Network(compiler.load_bittner(name))
B = 1 << 20
sim = engine.Simulator(net, B, seed=1)
sim.rand_state()
tgt = np.arange(7, dtype=np.int32)
sim.ssd(100, 0.01, tgt)
t = timed(lambda: sim.ssd(2000, 0.01, tgt))
t2 = timed(lambda: sim.rollout(2000))
t3 = timed(lambda: sim.rollout(20, sync=True))
print(f"{abi.LIB_PATH.name} {name}: ssd {B*2000/t:.4g} it/s | async rollout {B*2000/t2:.4g} upd/s | sync {B*20*net.n/t3:.4g} node-upd/s")
