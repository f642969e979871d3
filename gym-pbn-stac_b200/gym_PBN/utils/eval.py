"""Steady-state-distribution (SSD) estimation on the GPU.

Drop-in for the reference's `compute_ssd_hist` (gym_PBN/utils/eval.py:20-72) and its worker `_ssd_run` (:76-103):
`resets` independent chains, each `iters // resets` iterations of  histogram(target-gene pattern) -> flip every gene
w.p. bit_flip_prob -> env.step(0).  Here every chain is one GPU thread of the K3 kernel (pbn_ssd), the per-chain
float32 histograms + np.mean of the reference become one exact uint64 histogram, and with torch.distributed
initialised the chains are sharded over ranks and the histogram is all-reduced (NCCL).
"""
import itertools

import numpy as np
import torch

from gym_PBN.b200 import abi
from gym_PBN.b200 import dist as pdist
from gym_PBN.b200 import engine

MAX_ITERS_PER_LAUNCH = 1 << 22  # keeps a block's 32-bit shared-memory bucket counts far from overflow


def _run_chains(net, sim, iters, p, tgt, env_image, hist):
    left = int(iters)
    while left > 0:
        chunk = min(left, MAX_ITERS_PER_LAUNCH)
        sim.ssd(chunk, p, tgt, env=env_image, hist=hist)
        left -= chunk
    return hist


_HOST_SIMS = {}  # (network handle, chains) -> Simulator: repeated host-side estimates reuse the device buffers


def ssd_histogram_host(net, start_states, iters, bit_flip_prob, tgt_nodes, env_image=None, seed=0, env0=0, distributed=False):
    """Visit histogram of `chains` chains x `iters` iterations, HOST in / HOST out.

    start_states: uint8/bool [chains][N] host array or (pinned) CPU tensor — what env.render() hands out — or the same
    states BIT-PACKED as int32/uint32 planes [W32][chains] (node i = bit i & 31 of plane i >> 5: 16 bytes per chain for a
    100-node network instead of 100, which is what crosses PCIe).  Returns a NumPy int64 [2^g] histogram (summed over ranks
    when distributed=True)."""
    st = start_states if torch.is_tensor(start_states) else torch.from_numpy(np.ascontiguousarray(start_states))
    packed = st.dtype in (torch.int32, torch.uint32) and st.dim() == 2 and st.shape[0] == net.w32
    chains = st.shape[1] if packed else st.shape[0]
    key = (net.handle.value, chains)
    sim = _HOST_SIMS.get(key)
    if sim is None:
        if len(_HOST_SIMS) >= 4:
            _HOST_SIMS.clear()
        sim = _HOST_SIMS[key] = engine.Simulator(net, chains)
    sim.reseed(seed)
    sim.env0 = int(env0)
    if packed:
        sim.state.copy_(st.view(torch.int32) if st.dtype != torch.int32 else st, non_blocking=True)
    else:
        sim.set_state(st.to(torch.uint8).to(net.device, non_blocking=True))
    tgt = np.ascontiguousarray(tgt_nodes, np.int32)
    hist = torch.zeros(1 << len(tgt), dtype=torch.int64, device=net.device)
    _run_chains(net, sim, iters, bit_flip_prob, tgt, env_image, hist)
    if distributed:
        pdist.allreduce_sum_(hist)
    return hist.cpu().numpy()


def pack_states(bits):
    """uint8/bool [chains][N] -> int32 planes [W32][chains] (host NumPy): the packed form ssd_histogram_host accepts."""
    b = np.ascontiguousarray(np.asarray(bits, dtype=np.uint8))
    chains, n = b.shape
    w32 = (n + 31) // 32
    pad = np.zeros((chains, w32 * 32), np.uint8)
    pad[:, :n] = b
    words = np.packbits(pad.reshape(chains, w32, 32), axis=2, bitorder="little").view(np.uint32).reshape(chains, w32)
    return np.ascontiguousarray(words.T).view(np.int32)


def ssd_histogram(env, iters, chains, bit_flip_prob=0.01, seed=None, distributed=True):
    """Device-side estimate for one of our envs: chains start from env.reset() states (or uniform random states when the
    env has no attractor list), sharded over ranks by global chain id.  Returns (int64 device histogram, total iterations)."""
    net = env.network
    start, stop = pdist.shard_range(chains, align=32) if distributed else (0, chains)
    local = stop - start
    seed = env._next_seed() if seed is None else seed
    sim = engine.Simulator(net, max(local, 1), seed=seed, env0=start)
    image = env.env_image
    if local > 0:
        if image is not None and image.n_att >= 2:
            sim.env_reset(image)       # env.reset() per chain (eval.py:78)
        else:
            sim.rand_state()           # no attractors known: Graph.genRandState
    tgt = np.asarray(env.target_node_indices, np.int32)
    hist = torch.zeros(1 << len(tgt), dtype=torch.int64, device=net.device)
    per_chain = int(iters)
    if local > 0:
        _run_chains(net, sim, per_chain, bit_flip_prob, tgt, image, hist)
    if distributed:
        pdist.allreduce_sum_(hist)
    return hist, per_chain * chains


def _policy_actions(model, obs):
    """Batched actions for uint8 observations [B][N] on the device.  A model with `predict_batch(obs)` is called once on
    the device tensor; otherwise the reference protocol `model.predict(state, target, deterministic=True)` (target = state,
    utils/eval.py:81-82,98) is called per chain on host arrays."""
    if hasattr(model, "predict_batch"):
        a = model.predict_batch(obs)
        return torch.as_tensor(a).to(obs.device, dtype=torch.int32).reshape(obs.shape[0], -1)
    host = obs.cpu().numpy()
    acts = []
    for row in host:
        a = model.predict(row, row, deterministic=True)
        if type(a) == tuple:
            a = a[0]
        acts.append(np.asarray(a).reshape(-1))
    return torch.as_tensor(np.stack(acts).astype(np.int32)).to(obs.device)


def ssd_histogram_policy(env, model, iters, chains, seed=None, distributed=True):
    """The `model` branch of _ssd_run (utils/eval.py:97-101): per iteration  histogram -> action = model.predict(state)
    -> env.step(action), no perturbation, no reset inside a chain.  All chains advance in lockstep through the fused
    vector step; the histogram update is its own small kernel (pbn_bucket_hist)."""
    from gym_PBN.b200.vector_env import PBNVectorEnv

    start, stop = pdist.shard_range(chains) if distributed else (0, chains)
    seed = env._next_seed() if seed is None else seed
    tgt = np.asarray(env.target_node_indices, np.int32)
    hist = torch.zeros(1 << len(tgt), dtype=torch.int64, device=env.network.device)
    if stop > start:
        vec = PBNVectorEnv(env, stop - start, seed=seed, autoreset=False)
        vec.sim.env0 = start
        obs, _ = vec.reset()
        for _ in range(int(iters)):
            engine.bucket_hist(vec.sim, tgt, hist)
            obs, *_ = vec.step(_policy_actions(model, obs))
    if distributed:
        pdist.allreduce_sum_(hist)
    return hist, int(iters) * chains


def compute_ssd_hist(env, model=None, iters=1_200_000, resets=300, bit_flip_prob=0.01, multiprocess=True, seed=None):
    """Reference signature (utils/eval.py:20-27).  Returns (DataFrame indexed '0..0'..'1..1' MSB-first with column
    "Value", figure-or-None).  `multiprocess` is accepted for compatibility; parallelism here is the GPU's."""
    SSD_N, SSD_RESETS = int(iters), int(resets)
    assert bit_flip_prob >= 0 and bit_flip_prob <= 1, "Invalid Bit Flip Probability value."
    assert SSD_RESETS > 0, "Invalid resets value."
    assert SSD_N > 0, "Invalid iterations value."
    assert SSD_N // SSD_RESETS, "Resets does not divide the iterations."
    g = len(env.target_nodes)
    if model is not None:
        hist, total = ssd_histogram_policy(env, model, SSD_N // SSD_RESETS, SSD_RESETS, seed=seed)
    else:
        hist, total = ssd_histogram(env, SSD_N // SSD_RESETS, SSD_RESETS, bit_flip_prob, seed=seed)
    ssd = hist.cpu().numpy().astype(np.float64) / float(total)
    states = ["".join(str(b) for b in bits) for bits in itertools.product([0, 1], repeat=g)]
    try:
        import pandas as pd

        ret = pd.DataFrame(list(ssd), index=states, columns=["Value"])
    except Exception:  # pandas is optional plumbing
        ret = {"index": states, "Value": ssd}
    return ret, visualize_ssd(ret, getattr(env, "name", None))


def eval_increase(env, model, original_ssd=None, iters=1_200_000, resets=300, bit_flip_prob=0.01):
    """Total increase of the favourable target-gene patterns (env.target_node_values) in the SSD under `model` versus the
    uncontrolled SSD (reference: utils/eval.py:106-136; there the subtraction is applied to the (frame, figure) tuples
    that compute_ssd_hist returns and cannot run — here it is applied to the frames)."""
    if original_ssd is None:
        original_ssd = compute_ssd_hist(env, iters=iters, resets=resets, bit_flip_prob=bit_flip_prob)
    model_ssd = compute_ssd_hist(env, model, iters=iters, resets=resets, bit_flip_prob=bit_flip_prob)
    frame = lambda x: x[0] if isinstance(x, tuple) else x  # noqa: E731
    states = ["".join(str(int(b)) for b in state) for state in env.target_node_values]
    delta = frame(model_ssd) - frame(original_ssd)
    return float(delta.loc[states, "Value"].sum())


def _step_limit(env):
    """max_episode_steps of an env made through the registry: the spec's value, else the TimeLimit wrapper's own counter
    (gymnasium keeps it in `_max_episode_steps`, which Wrapper.__getattr__ hides, so the wrappers are walked; the in-repo
    stand-in calls it `_max`)."""
    e = env
    while e is not None:
        spec = e.__dict__.get("spec") if hasattr(e, "__dict__") else None
        if spec is not None and getattr(spec, "max_episode_steps", None):
            return int(spec.max_episode_steps)
        for name in ("_max_episode_steps", "_max"):
            v = e.__dict__.get(name) if hasattr(e, "__dict__") else None
            if v:
                return int(v)
        e = e.__dict__.get("env") if hasattr(e, "__dict__") else None
    return None


def eval_winrate(env, model, max_states=200_000, max_episode_steps=None, seed=0):
    """(win rate, mean interactions, mean time steps) of `model` started from every non-target state, as the reference's
    eval_winrate (utils/eval.py:160-197) computes them: states in itertools.product([0, 1], repeat=N) order (node 0 most
    significant), target states skipped, enumeration stopped after index max_states + 1; an episode ends when the env
    reports terminated (a win) or truncated; the start state is written with env.set (reset(options=) re-draws it, pbn_env.py:195-206); "time steps" sums info["interval"] (1 for envs without intervals).
    The reference version cannot return (a debugging `raise ValueError` fires on the first win, :181); this one runs all
    start states in lockstep on the GPU.  Truncation comes from the registration's max_episode_steps (the TimeLimit
    wrapper gym.make adds), or `max_episode_steps` here.  `model.predict_batch(obs uint8 [B][N]) -> actions [B][width]`
    is called once per step when it exists, else `model.predict(obs_row, obs_row, deterministic=True)` per env."""
    from gym_PBN.b200.vector_env import PBNVectorEnv

    limit = max_episode_steps if max_episode_steps is not None else _step_limit(env)
    if not limit:
        raise ValueError("eval_winrate needs a step limit: make the env through gym_PBN.make (registered max_episode_steps) "
                         "or pass max_episode_steps")
    core = getattr(env, "unwrapped", env)
    n = core.observation_space.n
    if n > 40:
        raise ValueError("eval_winrate enumerates all 2^N states")
    count = min(1 << n, int(max_states) + 2)
    idx = np.arange(count, dtype=np.int64)
    bits = ((idx[:, None] >> np.arange(n - 1, -1, -1)) & 1).astype(np.uint8)
    targets = {tuple(int(v) for v in t) for t in (getattr(core, "target", None) or core.target_nodes)}  # PBNEnv keeps them as target_nodes
    keep = np.array([tuple(int(v) for v in row) not in targets for row in bits]) if targets else np.ones(count, bool)
    starts = bits[keep]
    iters = len(starts)
    if iters == 0:
        return float("nan"), float("nan"), float("nan")
    vec = PBNVectorEnv(core, iters, seed=seed, autoreset=False)
    vec.reset()
    vec.sim.set_state(torch.as_tensor(starts, device=vec.device))
    vec.sim.n_steps.zero_()
    dev = vec.device
    alive = torch.ones(iters, dtype=torch.bool, device=dev)
    won = torch.zeros(iters, dtype=torch.bool, device=dev)
    interactions = torch.zeros(iters, dtype=torch.int64, device=dev)
    timesteps = torch.zeros(iters, dtype=torch.int64, device=dev)
    interval_col = {abi.ENV_PBN_SD: 1, abi.ENV_PBCN_SD: 0}.get(vec.image.kind)
    obs = vec.sim.unpack()
    for t in range(int(limit)):
        actions = _policy_actions(model, obs)
        obs, _r, terminated, _trunc, _info = vec.step(actions)
        if interval_col is None:
            step_len = 1
        else:  # what the single env reports as info["interval"]: the last loop index for PBN-sampled-data (sampled_data.py:84)
            step_len = actions[:, interval_col].to(torch.int64) - (1 if vec.image.kind == abi.ENV_PBN_SD else 0)
        interactions += alive
        timesteps += alive * step_len
        won |= alive & terminated.bool()
        alive &= ~terminated.bool()
        if not bool(alive.any()):
            break
    return float(won.sum().item()) / iters, float(interactions.double().mean().item()), float(timesteps.double().mean().item())


def visualize_ssd(ssd_frame, env_name):
    """Bar chart of the estimate when plotly is installed (reference: utils/eval.py:139-157); None otherwise."""
    try:
        import plotly.express as px

        fig = px.bar(ssd_frame, x=list(ssd_frame.index), y="Value", title=f"SSD for {env_name}")
        return fig
    except Exception:
        return None


def total_variation(p, q):
    p, q = np.asarray(p, np.float64), np.asarray(q, np.float64)
    return 0.5 * float(np.abs(p / p.sum() - q / q.sum()).sum())
