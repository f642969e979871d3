"""CPU, world_size 2 over gloo: the N>1 host logic — sharding by global env id and the SUM all-reduce of the SSD
histogram / episode statistics.  (Each rank's histogram here comes from the oracle in Philox mode, which is the
stream the CUDA kernels use: two half-ranges must add up to the one-rank result.)"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, out_dir):
    for p in (ROOT / "gym-pbn-stac_b200", ROOT / "oracle"):
        sys.path.insert(0, str(p))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as orc
    from gym_PBN.b200 import dist as pdist

    sets, ids = orc.load_bittner("28_15_median")
    net = orc.net_from_predictor_sets(sets, ids)
    total, iters, seed = 600, 50, 11
    start, stop = pdist.shard_range(total, align=32)
    assert start % 32 == 0
    st = orc.rand_state(net, stop - start, orc.Draws(seed=seed, epoch=0), env0=start)
    h = orc.ssd(net, None, st, iters, 0.01, np.arange(5, dtype=np.int32), orc.Draws(seed=seed, epoch=1), env0=start)
    t = torch.from_numpy(h.astype(np.int64))
    pdist.allreduce_sum_(t)
    stats = pdist.EpisodeStats("cpu")
    stats.v += torch.tensor([1, 10 * (rank + 1), 5, 1, 0, stop - start, 0, 0])
    red = stats.reduced()
    if rank == 0:
        np.save(os.path.join(out_dir, "hist.npy"), t.numpy())
        np.save(os.path.join(out_dir, "stats.npy"), np.array([red["episodes"], red["return_sum"], red["env_steps"]]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_histogram_equals_one_rank(tmp_path):
    import oracle as orc

    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "hist.npy")
    sets, ids = orc.load_bittner("28_15_median")
    net = orc.net_from_predictor_sets(sets, ids)
    st = orc.rand_state(net, 600, orc.Draws(seed=11, epoch=0))
    want = orc.ssd(net, None, st, 50, 0.01, np.arange(5, dtype=np.int32), orc.Draws(seed=11, epoch=1))
    assert np.array_equal(got.astype(np.uint64), want)
    assert np.load(tmp_path / "stats.npy").tolist() == [2, 30, 600]
