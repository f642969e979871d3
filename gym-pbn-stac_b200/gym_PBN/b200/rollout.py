"""Device-resident rollout storage for agents trained on PBNVectorEnv (agent integration).

The reference is used with single-env agents (DDQN/BDQ, stable-baselines style loops: README.md:12,21) that copy every
transition through host memory.  Here a rollout of T steps over B lockstep envs stays on the GPU: observations as uint8
[T+1][B][N] (or packed int32 planes), actions, rewards, terminated/truncated flags, and the quantities a policy-gradient
or a TD learner needs from them (discounted returns, GAE advantages, flat transition minibatches), all as torch ops on
the device.  Plumbing, not a kernel: the env step itself is the fused CUDA launch of PBNVectorEnv.step.
"""
import torch


class RolloutBuffer:
    def __init__(self, vec_env, horizon, gamma=0.99, gae_lambda=0.95, action_width=None):
        self.env, self.T, self.B = vec_env, int(horizon), vec_env.num_envs
        self.gamma, self.lam = float(gamma), float(gae_lambda)
        dev = vec_env.device
        n = vec_env.n
        w = action_width or vec_env.action_width
        self.obs = torch.zeros((self.T + 1, self.B, n), dtype=torch.uint8, device=dev)
        self.actions = torch.zeros((self.T, self.B, w), dtype=torch.int32, device=dev)
        self.rewards = torch.zeros((self.T, self.B), dtype=torch.float32, device=dev)
        self.terminated = torch.zeros((self.T, self.B), dtype=torch.bool, device=dev)
        self.truncated = torch.zeros((self.T, self.B), dtype=torch.bool, device=dev)
        self.values = torch.zeros((self.T + 1, self.B), dtype=torch.float32, device=dev)
        self.log_probs = torch.zeros((self.T, self.B), dtype=torch.float32, device=dev)
        self.t = 0

    def start(self, obs):
        self.obs[0].copy_(obs)
        self.t = 0

    def add(self, actions, next_obs, reward, terminated, truncated, value=None, log_prob=None):
        """Stores transition t; `next_obs` is what PBNVectorEnv.step returned (for auto-reset envs: the new episode's start)."""
        t = self.t
        if t >= self.T:
            raise IndexError("rollout buffer is full")
        self.actions[t].copy_(actions.reshape(self.B, -1))
        self.obs[t + 1].copy_(next_obs)
        self.rewards[t].copy_(reward)
        self.terminated[t].copy_(terminated.bool())
        self.truncated[t].copy_(truncated.bool())
        if value is not None:
            self.values[t].copy_(value)
        if log_prob is not None:
            self.log_probs[t].copy_(log_prob)
        self.t = t + 1

    @property
    def done(self):
        return self.terminated | self.truncated

    def returns(self, last_value=None):
        """Discounted return-to-go [T][B]; episodes end at terminated/truncated, bootstrapped with `last_value` [B] at the
        end of the buffer (0 if None)."""
        return discounted_returns(self.rewards[: self.t], self.done[: self.t], self.gamma, last_value)

    def gae(self, last_value):
        """(advantages, returns) [T][B] by generalised advantage estimation from `values` stored with add().  An episode
        end of either kind stops the recursion and does not bootstrap: with auto-reset the stored next observation
        already belongs to the next episode (the step's own last observation is in info["final_obs_packed"])."""
        self.values[self.t].copy_(last_value)
        return gae_advantages(self.rewards[: self.t], self.values[: self.t + 1], self.terminated[: self.t],
                              self.truncated[: self.t], self.gamma, self.lam)

    def transitions(self):
        """Flat (obs, action, reward, next_obs, terminated) over the T*B stored transitions (views, no copy)."""
        t = self.t
        return (self.obs[:t].reshape(t * self.B, -1), self.actions[:t].reshape(t * self.B, -1),
                self.rewards[:t].reshape(-1), self.obs[1:t + 1].reshape(t * self.B, -1), self.terminated[:t].reshape(-1))

    def collect(self, policy, obs=None):
        """Runs the env for `horizon` steps under `policy(obs uint8 [B][N]) -> (actions int32 [B][w], log_prob [B] or None,
        value [B] or None)`; returns the observation after the last step."""
        if obs is None:
            obs, _ = self.env.reset()
        self.start(obs)
        for _ in range(self.T):
            with torch.no_grad():
                actions, log_prob, value = policy(self.obs[self.t])
            obs, reward, terminated, truncated, _ = self.env.step(actions)
            self.add(actions, obs, reward, terminated, truncated, value, log_prob)
        return self.obs[self.t]


def discounted_returns(rewards, done, gamma, last_value=None):
    T = rewards.shape[0]
    out = torch.empty_like(rewards, dtype=torch.float32)
    running = torch.zeros_like(rewards[0], dtype=torch.float32) if last_value is None else last_value.to(torch.float32).clone()
    for t in range(T - 1, -1, -1):
        running = rewards[t].to(torch.float32) + gamma * running * (~done[t]).to(torch.float32)
        out[t] = running
    return out


def gae_advantages(rewards, values, terminated, truncated, gamma, lam):
    T = rewards.shape[0]
    adv = torch.empty_like(rewards, dtype=torch.float32)
    running = torch.zeros_like(rewards[0], dtype=torch.float32)
    for t in range(T - 1, -1, -1):
        cont = (~(terminated[t] | truncated[t])).to(torch.float32)
        delta = rewards[t].to(torch.float32) + gamma * values[t + 1] * cont - values[t]
        running = delta + gamma * lam * cont * running
        adv[t] = running
    return adv, adv + values[:T]
