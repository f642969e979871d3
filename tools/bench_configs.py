"""Throughput of the other BASELINE.json configurations (parity cases; bench.py stays on the headline workload).
Writes one JSON object per configuration to stdout; run on a B200:  python tools/bench_configs.py > profiles/rNN_configs.json"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gym-pbn-stac_b200"))
sys.path.insert(0, str(ROOT / "tests"))
from gym_PBN.b200 import abi, compiler, engine  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / reps


def cube_fixture(n, rng, n_att=4, care=4):
    atts = []
    for _ in range(n_att):
        c = ["*"] * n
        for i in rng.choice(n, size=care, replace=False):
            c[i] = int(rng.integers(0, 2))
        atts.append([tuple(c)])
    return atts


def main():
    out = []
    rng = np.random.default_rng(0)
    torch.manual_seed(0)
    # configs[1]: Bittner-28 PBN-target-v0, 65 536 lockstep envs, one env.step per launch (+ masked reset launch)
    net = engine.Network(compiler.load_bittner("28_15_median"))
    B = 65536
    from gym_PBN.b200 import attractors as att_tools

    atts = att_tools.exact_attractor_cubes(net)  # the real attractors of the 28-gene network (120 + 49152 states)
    env = engine.EnvImage(net, abi.ENV_TARGET, attractors=atts, horizon=100, max_inner=4096)
    sim = engine.Simulator(net, B, seed=1)
    sim.env_reset(env)
    acts = torch.randint(0, 29, (B, 1), dtype=torch.int32, device="cuda")

    def step28():
        sim.env_step(env, acts)
        sim.env_reset(env, mask=(sim.terminated | sim.truncated))

    t = timed(step28, reps=50)
    out.append({"config": "Bittner-28 PBN-target-v0 with its exact attractors, 65536 lockstep envs, step+auto-reset (2 launches)", "env_steps_per_s": B / t,
                "ms_per_step": t * 1e3, "mean_inner_updates": float(sim.inner.float().mean()), "max_inner": int(sim.inner.max())})
    # the same env at 2^20 envs: enough work to hide the serial tail of the slowest env
    BB = 1 << 20
    simb = engine.Simulator(net, BB, seed=1)
    simb.env_reset(env)
    actsb = torch.randint(0, 29, (BB, 1), dtype=torch.int32, device="cuda")

    def step28b():
        simb.env_step(env, actsb)
        simb.env_reset(env, mask=(simb.terminated | simb.truncated))

    t = timed(step28b, reps=10)
    inner = float(simb.inner.float().mean())
    out.append({"config": "Bittner-28 PBN-target-v0 with its exact attractors, 1048576 envs, step+auto-reset", "env_steps_per_s": BB / t,
                "node_updates_per_s": BB * inner / t, "ms_per_step": t * 1e3, "mean_inner_updates": inner})
    # same, all-attracting (exactly one update per env.step)
    env1 = engine.EnvImage(net, abi.ENV_TARGET, attractors=[[("*",) * 28], [("*",) * 28]], horizon=100, max_inner=1)
    sim.env_reset(env1)
    t = timed(lambda: sim.env_step(env1, acts), reps=50)
    out.append({"config": "Bittner-28 PBN-target-v0, 65536 envs, all-attracting, 1 launch/step", "env_steps_per_s": B / t, "ms_per_step": t * 1e3})
    # the same through the public PBNVectorEnv.step (fused launch + unpack), wall clock, device-resident actions
    import time

    import gym_PBN
    from gym_PBN.b200.vector_env import PBNVectorEnv

    genv = gym_PBN.make("gym-PBN/Bittner-28-v0", all_attractors=[[("*",) * 28], [("*",) * 28]], max_inner_steps=1)
    for obs_mode, graph in (("bits", False), ("packed", False), ("bits", True), ("packed", True)):
        vec = PBNVectorEnv(genv, 65536, seed=1, obs=obs_mode, cuda_graph=graph)
        vec.reset()
        for _ in range(20):
            vec.step(acts)
        a_in = acts
        if graph:  # the policy writes straight into the buffer the captured graph reads
            a_in = vec.action_buffer
            a_in.copy_(acts)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(200):
            vec.step(a_in)
        torch.cuda.synchronize()
        t = (time.perf_counter() - t0) / 200
        out.append({"config": f"PBNVectorEnv.step wall clock, Bittner-28, 65536 envs, all-attracting, obs={obs_mode}, cuda_graph={graph}",
                    "env_steps_per_s": 65536 / t, "us_per_step": t * 1e6})
    # configs[3]: Bittner-200 target_multi, attractor path, 131 072 envs per GPU
    net = engine.Network(compiler.load_bittner("200_5_kmeans"))
    B = 131072
    atts = cube_fixture(net.n, rng, care=4)
    env = engine.EnvImage(net, abi.ENV_MULTI, attractors=atts, horizon=100, max_inner=4096, dedup=True)
    sim = engine.Simulator(net, B, seed=2)
    sim.env_reset(env)
    acts = torch.randint(0, net.n + 1, (B, 3), dtype=torch.int32, device="cuda")

    def step200():
        sim.env_step(env, acts)
        sim.env_reset(env, mask=(sim.terminated | sim.truncated))

    t = timed(step200, reps=20)
    inner = float(sim.inner.float().mean())
    out.append({"config": "Bittner-200 (199 nodes) PBN-target_multi-v0, 131072 envs, K=3, cap 4096", "env_steps_per_s": B / t,
                "node_updates_per_s": B * inner / t, "ms_per_step": t * 1e3, "mean_inner_updates": inner,
                "cap_hits": int((sim.inner >= 4096).sum())})
    # configs[4]: synthetic PBCN N=1024, sampled-data, interval ~ U{1..64}
    from golden_util import synthetic_pbcn

    data = synthetic_pbcn()
    net = engine.Network(compiler.compile_pbn_data(data))
    B = 262144
    targets = [tuple(int(v) for v in rng.integers(0, 2, 1024)) for _ in range(4)]
    for cw in (False, True):
        env = engine.EnvImage(net, abi.ENV_PBCN_SD, attractors=[[t] for t in targets], targets=targets[:1], n_control=8,
                              control_write=cw)
        sim = engine.Simulator(net, B, seed=3)
        sim.rand_state()
        acts = torch.cat([torch.randint(1, 65, (B, 1)), torch.randint(0, 2, (B, 8))], 1).to(torch.int32).cuda()
        t = timed(lambda: sim.env_step(env, acts), reps=10)
        upd = float(acts[:, 0].float().sum())
        out.append({"config": f"synthetic PBCN N=1024 sampled-data, 262144 envs, interval~U(1..64), control={'write' if cw else 'stac'}",
                    "env_steps_per_s": B / t, "node_updates_per_s": upd / t, "ms_per_step": t * 1e3})
    # Bittner-100 rollouts (context for the headline SSD number)
    net = engine.Network(compiler.load_bittner("100_5_kmeans"))
    B = 1 << 20
    sim = engine.Simulator(net, B, seed=4)
    sim.rand_state()
    t = timed(lambda: sim.rollout(2000), reps=3)
    out.append({"config": "Bittner-100 async rollout, 2^20 envs x 2000 updates/launch", "node_updates_per_s": B * 2000 / t})
    t = timed(lambda: sim.rollout(20, sync=True), reps=3)
    out.append({"config": "Bittner-100 sync rollout, 2^20 envs x 20 steps/launch", "env_steps_per_s": B * 20 / t,
                "node_updates_per_s": B * 20 * 100 / t})
    t = timed(lambda: sim.rollout(50, sync="sliced"), reps=3)
    out.append({"config": "Bittner-100 sync rollout, bit-sliced kernel (32 envs per word), 2^20 envs x 50 steps/launch",
                "env_steps_per_s": B * 50 / t, "node_updates_per_s": B * 50 * 100 / t})
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
