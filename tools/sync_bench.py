import sys; sys.path.insert(0,"gym-pbn-stac_b200")
import torch, numpy as np
from gym_PBN.b200 import compiler, engine
def timed(fn, reps=3):
    best=1e9
    for _ in range(reps):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best=min(best,a.elapsed_time(b)*1e-3)
    return best
for name in ("100_5_kmeans","200_5_kmeans"):
    net=engine.Network(compiler.load_bittner(name)); B=1<<20
    sim=engine.Simulator(net,B,seed=1); sim.rand_state(); sim.rollout(2,sync="sliced")
    t1=timed(lambda: sim.rollout(50,sync="sliced")); t0=timed(lambda: sim.rollout(20,sync=True))
    print(name,"sync sliced: %.4g env-steps/s (%.4g node-upd/s) | per-env sync: %.4g env-steps/s" % (B*50/t1, B*50*net.n/t1, B*20/t0))
