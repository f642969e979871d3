"""PBNVectorEnv — `num_envs` lockstep copies of one gym_PBN env on one GPU, everything resident on the device.

The reference has no vectorised env (one Python object per env).  Here one CUDA launch performs
`env.step` for every env — interventions, update(s) until attracting (capped), reward / terminated / truncated — and, in
the same launch (pbn_vec_step), the episode bookkeeping, the statistics and the reset of the envs that finished.  Observations, rewards and flags are torch tensors on the device, owned by the
env and overwritten by the next `step` (clone them to keep them).

Multi-GPU: one process per GPU; pass `global_num_envs` and the env ids are sharded over ranks
(gym_PBN.b200.dist.shard_range).  Every env's Philox stream is keyed by its GLOBAL id, so an N-rank job reproduces
the 1-rank job bit for bit; `stats.reduced()` all-reduces the episode statistics.
"""
import numpy as np
import torch

from . import abi, dist as pdist, engine

_KIND_OF = {"PBNEnv": abi.ENV_PBN, "PBCNEnv": abi.ENV_PBCN, "PBNSampledDataEnv": abi.ENV_PBN_SD,
            "PBCNSampledDataEnv": abi.ENV_PBCN_SD}


def _family(env):
    from gym_PBN.envs.pbn_env import PBNEnv
    from gym_PBN.envs.pbn_target import PBNTargetEnv
    from gym_PBN.envs.pbn_target_multi import PBNTargetMultiEnv

    if "SelfTriggering" in type(env).__name__:
        return "st"
    if isinstance(env, PBNTargetMultiEnv):
        return "multi"
    if isinstance(env, PBNTargetEnv):
        return "target"
    if isinstance(env, PBNEnv):
        return "pbn"
    raise TypeError(f"cannot vectorise {type(env).__name__}")


class PBNVectorEnv:
    def __init__(self, env, num_envs=None, seed=0, autoreset=True, global_num_envs=None, obs="bits", dedup=True,
                 max_inner_steps=None, force=False, action_slots=3, curriculum=False, sample_pair=None, cuda_graph=False):
        env = getattr(env, "unwrapped", env)
        self.env = env
        self.family = _family(env)
        self.network = env.network
        self.device = self.network.device
        if global_num_envs is not None:
            start, stop = pdist.shard_range(global_num_envs)
            num_envs, env0 = stop - start, start
        else:
            env0 = 0
        if not num_envs or num_envs < 1:
            raise ValueError("num_envs must be >= 1")
        self.num_envs, self.env0 = int(num_envs), int(env0)
        self.autoreset, self.obs_mode = bool(autoreset), obs
        self.n = self.network.n
        self.sim = engine.Simulator(self.network, self.num_envs, seed=seed, env0=self.env0)
        self.single_observation_space = env.observation_space
        self.single_action_space = env.action_space
        self.max_inner_steps = int(max_inner_steps if max_inner_steps is not None else getattr(env, "max_inner_steps", 1))
        self.action_slots = action_slots
        if self.family == "st":
            # self-triggering macro steps (envs/self_triggering.py): one launch of the macro-step kernel, float64 discounted
            # rewards; bookkeeping and the masked reset follow as separate small launches (no fused epilogue for float returns)
            self.image = env._st_image()
            self.action_width = 2 if self.image.kind == abi.ENV_PBN_ST else 1 + env.PBN.M  # (primitive, prob) | (prob, control..)
            self.horizon = 0
        elif self.family == "pbn":
            kind = next(_KIND_OF[c.__name__] for c in type(env).__mro__ if c.__name__ in _KIND_OF)  # most derived class first
            self.image = engine.EnvImage(
                self.network, kind, attractors=[sorted(a) for a in env.all_attractors], targets=sorted(env.target_nodes),
                n_control=getattr(env.PBN, "M", 0), control_write=getattr(env.PBN, "control_mode", "stac") == "write",
                successful_reward=env.successful_reward, wrong_attractor_cost=env.wrong_attractor_cost)
            self.action_width = {abi.ENV_PBN: 1, abi.ENV_PBCN: 1, abi.ENV_PBN_SD: 2}.get(kind, 1 + getattr(env.PBN, "M", 0))
            self.horizon = 0
        else:
            atts = env.all_attractors or [[("*",) * self.n], [("*",) * self.n]]  # all-attracting fixture
            self.image = engine.EnvImage(
                self.network, abi.ENV_TARGET if self.family == "target" else abi.ENV_MULTI, attractors=atts,
                horizon=env.horizon, max_inner=self.max_inner_steps, force=force, dedup=dedup)
            self.action_width = 1 if self.family == "target" else action_slots
            self.horizon = env.horizon
        # curriculum of PBNTargetMultiEnv (pbn_target_multi.py:159-181, 232-235) on the device: every env keeps its own
        # probability row (= B independent env objects of the reference); an episode's end applies rework_probas(episode
        # length) to the row inside the step launch, and the reset draws the (state, target) attractor ids from it
        self.probabilities = self.pair_ids = None
        self.sample_pair = bool(getattr(env, "sample_pair", False) if sample_pair is None else sample_pair)
        if curriculum:
            n_att = self.image.n_att
            if self.family != "multi" or not 2 <= n_att <= 64:
                raise ValueError("curriculum=True needs a PBNTargetMultiEnv with 2..64 attractors")
            self.probabilities = torch.full((self.num_envs, n_att), 1.0 / n_att, dtype=torch.float64, device=self.device)
            self.pair_ids = torch.zeros((self.num_envs, 2), dtype=torch.int32, device=self.device)
        # cuda_graph=True: from its second call on, step() replays ONE captured CUDA graph (the step launches, the observation
        # unpack, the epoch increment) instead of issuing the launches from Python — the Philox epoch then lives in device
        # memory (PbnDraws.epoch_dev).  Same results, bit for bit; actions are copied into a buffer the graph reads.
        self.cuda_graph = bool(cuda_graph) and self.family != "st"
        self._graph = None
        self.stats = pdist.EpisodeStats(self.device)
        self.ep_return = torch.zeros(self.num_envs, dtype=torch.float64 if self.family == "st" else torch.int64,
                                     device=self.device)
        self._return_sum = torch.zeros(1, dtype=torch.float64, device=self.device)  # "st": stats.v[1] stays 0
        self.ep_len = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)
        self.final_obs = torch.zeros_like(self.sim.state)
        self._needs_reset = True
        self._warm = False
        self._obs_bits = None

    # ---- observations -----------------------------------------------------------------------------------------
    def _obs(self, planes):
        if self.obs_mode == "packed":
            return planes
        if self._obs_bits is None:  # owned by the env, overwritten by the next step (see the module docstring)
            self._obs_bits = torch.empty((self.num_envs, self.n), dtype=torch.uint8, device=self.device)
        return self.sim.unpack(planes, out=self._obs_bits)

    @property
    def state(self):
        """Live packed state planes int32 [W32][num_envs] (bit i&31 of word i>>5 = node i)."""
        return self.sim.state

    # ---- API --------------------------------------------------------------------------------------------------
    def reset(self, seed=None, options=None):
        if seed is not None:
            self.sim.reseed(seed)
        self.sim.env_reset(self.image, curriculum=self._curriculum())
        self.ep_return.zero_()
        self.ep_len.zero_()
        self._needs_reset = False
        info = {"target_attractor": self.sim.target_att, "target_state_packed": self.sim.target_state}
        return self._obs(self.sim.state), info

    def step(self, actions):
        if self._needs_reset:
            raise RuntimeError("call reset() before step()")
        if not torch.is_tensor(actions):
            actions = torch.as_tensor(np.asarray(actions))
        if actions.numel() != self.num_envs * self.action_width:
            raise ValueError(f"actions must have shape [{self.num_envs}, {self.action_width}]")
        sim = self.sim
        if self.family == "st":
            return self._step_self_triggering(actions)
        info = {"inner_steps": sim.inner, "packed_obs": sim.obs_state, "final_obs_packed": self.final_obs}
        if self.probabilities is not None:
            info["pair_ids"], info["probabilities"] = self.pair_ids, self.probabilities
        if self.cuda_graph and self._warm:
            return self._step_graph(actions), sim.reward, sim.terminated, sim.truncated, info
        # one fused launch: step + episode bookkeeping + statistics (+ reset of finished envs, own Philox epoch)
        sim.vec_step(self.image, actions, self.ep_return, self.ep_len, self.stats.v, final_obs=self.final_obs,
                     autoreset=self.autoreset, curriculum=self._curriculum())
        self._warm = True
        # after the launch obs_state holds the step's observation, or the NEW state for envs that were auto-reset
        return self._obs(sim.obs_state), sim.reward, sim.terminated, sim.truncated, info

    @property
    def action_buffer(self):
        """cuda_graph=True, after the graph exists: the int32 [num_envs][action_width] device tensor the captured step reads.
        A policy that writes its actions there and passes the same tensor to step() saves the copy."""
        return self._g_actions if self._graph is not None else None

    def _step_graph(self, actions):
        sim = self.sim
        actions = actions.to(self.device, dtype=torch.int32).reshape(self.num_envs, self.action_width)
        if self._graph is None:  # capture (the eager call before this one was the warm-up)
            self._g_actions = actions.clone().contiguous()
            self._g_epoch = torch.zeros(1, dtype=torch.int32, device=self.device)
            self._g_base = self._g_next = sim.epoch
            launches0, epoch0 = sim.launches, sim.epoch
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                sim.vec_step(self.image, self._g_actions, self.ep_return, self.ep_len, self.stats.v, final_obs=self.final_obs,
                             autoreset=self.autoreset, curriculum=self._curriculum(), epoch_dev=self._g_epoch, epoch_base=self._g_base)
                self._g_obs = self._obs(sim.obs_state)
                self._g_epoch += 2
            self._g_launches = sim.launches - launches0
            sim.launches, sim.epoch = launches0, epoch0  # capturing ran nothing
            self._graph = g
        elif actions.data_ptr() != self._g_actions.data_ptr():  # write into `action_buffer` to skip this copy
            self._g_actions.copy_(actions)
        if sim.epoch != self._g_next:  # epochs were consumed outside the graph (reset, load_state_dict): re-base the counter
            self._g_epoch.fill_(sim.epoch - self._g_base)
        self._graph.replay()
        sim.epoch += 2
        sim.launches += self._g_launches
        self._g_next = sim.epoch
        return self._g_obs

    def _curriculum(self):
        return None if self.probabilities is None else (self.probabilities, self.pair_ids, self.sample_pair)

    def _step_self_triggering(self, actions):
        """Macro step of every env (actions int32 [B][2] = (primitive, prob) for the PBN variant, [B][1+M] = (prob, control
        bits) for the PBCN variant); rewards are float64 [B], info["interval"] the primitive steps each env took.  ONE launch
        (pbn_vec_step): macro step + float64 return bookkeeping + statistics + reset of the finished envs, the same draw
        consumption as Simulator.env_step followed by a masked Simulator.env_reset (one epoch each)."""
        sim = self.sim
        if actions.dtype != torch.int32 or actions.device != self.device:
            actions = actions.to(self.device, dtype=torch.int32)
        sim.vec_step(self.image, actions.reshape(self.num_envs, self.action_width).contiguous(), self.ep_return, self.ep_len,
                     self.stats.v, final_obs=self.final_obs, autoreset=self.autoreset, return_sum_f64=self._return_sum)
        info = {"interval": sim.inner, "inner_steps": sim.inner, "packed_obs": sim.obs_state, "final_obs_packed": self.final_obs}
        return self._obs(sim.obs_state), sim.reward_f64, sim.terminated, sim.truncated, info

    @property
    def return_sum_f64(self):
        """Self-triggering envs: sum of the float64 returns of all finished episodes (device scalar)."""
        return self._return_sum[0]

    def step_host(self, actions_host):
        """Host in / host out convenience (pinned staging): NumPy actions -> NumPy (obs, reward, terminated, truncated)."""
        a = torch.as_tensor(np.ascontiguousarray(actions_host, dtype=np.int32))
        obs, r, te, tr, _ = self.step(a.pin_memory().to(self.device, non_blocking=True))
        return obs.cpu().numpy(), r.cpu().numpy(), te.cpu().numpy(), tr.cpu().numpy()

    def close(self):
        pass

    # ---- checkpoint -------------------------------------------------------------------------------------------
    def state_dict(self):
        s = self.sim
        return {"state": s.state.clone(), "n_steps": s.n_steps.clone(), "target_att": s.target_att.clone(),
                "target_state": s.target_state.clone(), "seed": s.seed, "epoch": s.epoch, "env0": s.env0,
                "ep_return": self.ep_return.clone(), "ep_len": self.ep_len.clone(), "stats": self.stats.v.clone(),
                "probabilities": None if self.probabilities is None else self.probabilities.clone(),
                "pair_ids": None if self.pair_ids is None else self.pair_ids.clone()}

    def load_state_dict(self, d):
        s = self.sim
        s.state.copy_(d["state"]); s.n_steps.copy_(d["n_steps"]); s.target_att.copy_(d["target_att"])
        s.target_state.copy_(d["target_state"])
        s.seed, s.epoch, s.env0 = int(d["seed"]), int(d["epoch"]), int(d["env0"])
        self.ep_return.copy_(d["ep_return"]); self.ep_len.copy_(d["ep_len"]); self.stats.v.copy_(d["stats"])
        if self.probabilities is not None and d.get("probabilities") is not None:
            self.probabilities.copy_(d["probabilities"]); self.pair_ids.copy_(d["pair_ids"])
        self._needs_reset = False


def make_vec(id, num_envs, vec_kwargs=None, **env_kwargs):
    import gym_PBN

    env = gym_PBN.make(id, **env_kwargs)
    return PBNVectorEnv(env, num_envs, **(vec_kwargs or {}))
