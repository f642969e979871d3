import sys; sys.path.insert(0,"gym-pbn-stac_b200")
import torch
from gym_PBN.b200 import compiler, engine
net=engine.Network(compiler.load_bittner("100_5_kmeans")); B=1<<20
sim=engine.Simulator(net,B,seed=1); sim.rand_state()
sim.rollout(10,sync="sliced"); torch.cuda.synchronize()
