"""Scratch probe: the headline SSD loop (Bittner-100, 7 targets, p = 0.01) at bench size, best of 3 launches."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gym-pbn-stac_b200"))
from gym_PBN.b200 import compiler, engine  # noqa: E402


def main():
    B = 1 << 20
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2400
    net = engine.Network(compiler.load_bittner("100_5_kmeans"))
    sim = engine.Simulator(net, B, seed=1)
    sim.rand_state()
    tgt = np.arange(7, dtype=np.int32)
    sim.ssd(100, 0.01, tgt)
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); sim.ssd(iters, 0.01, tgt); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    print(f"ssd B={B} iters={iters}: {B * iters / best:.4g} it/s  ({best * 1e3:.2f} ms)")


if __name__ == "__main__":
    main()
