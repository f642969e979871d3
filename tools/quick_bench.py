"""Scratch throughput probe (not the bench contract): rollout + SSD on Bittner-100, issue-peak microbenchmarks."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gym-pbn-stac_b200"))
from gym_PBN.b200 import compiler, engine  # noqa: E402


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    return best


def main():
    for kind, name in ((0, "int-alu ops/s"), (1, "philox blocks/s")):
        rate, ms = engine.issue_peak(kind, 4000)
        print(f"{name}: {rate:.4g}  ({ms:.2f} ms)")
    for name in ("100_5_kmeans", "28_15_median", "200_5_kmeans"):
        net = engine.Network(compiler.load_bittner(name))
        for B in (1 << 16, 1 << 18, 1 << 20):
            sim = engine.Simulator(net, B, seed=1)
            sim.rand_state()
            steps = 2000
            sim.rollout(10)
            t = timed(lambda: sim.rollout(steps))
            print(f"{name} rollout async B={B} steps={steps}: {B*steps/t:.4g} updates/s")
            t = timed(lambda: sim.rollout(20, sync=True))
            print(f"{name} rollout sync  B={B} steps=20: {B*20/t:.4g} env-steps/s, {B*20*net.n/t:.4g} node-updates/s")
            tgt = np.arange(7, dtype=np.int32)
            t = timed(lambda: sim.ssd(1000, 0.01, tgt))
            print(f"{name} ssd B={B} iters=1000: {B*1000/t:.4g} it/s")


if __name__ == "__main__":
    main()
