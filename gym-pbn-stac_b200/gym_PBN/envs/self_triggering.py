"""Self-triggering envs (reference: gym_PBN/envs/self_triggering.py): a macro action is (primitive, prob in 1..10);
after every primitive step it stops with probability prob/10 (or at i == T), inner rewards are discounted by gamma**i.
The whole macro step is ONE kernel launch (env kinds PBN_ENV_PBN_ST / PBN_ENV_PBCN_ST of pbn_env_step_f64): per primitive
step the draws are node index, node value, stop — the reference's order (self_triggering.py:64-80,165-181) — and the
discounted sum is accumulated in float64 from a gamma**i table computed by Python's own pow, so that under replayed draws
the reward equals the reference's bit for bit (tests/test_gpu_envs.py, tests/test_oracle_golden.py).
"""
import numpy as np
import torch

from gym_PBN.b200 import abi, engine
from gym_PBN.b200.gym_compat import spaces
from gym_PBN.utils import booleanize

from ._device import state_to_idx
from .pbcn_env import PBCNEnv
from .pbn_env import PBNEnv


def _macro_step(env, image, acts):
    """One launch of the self-triggering kernel for the single env -> (discounted reward, terminated, truncated, interval)."""
    queue = getattr(env, "_replays", None)
    replay = queue.pop(0) if queue else None
    sim = env.sim
    sim.env_step(image, torch.tensor([acts], dtype=torch.int32), replay=replay)
    out = (float(sim.reward_f64[0]), bool(sim.terminated[0]), bool(sim.truncated[0]), int(sim.inner[0]))
    env._last_state = sim.unpack()[0].cpu().numpy()
    return out


class PBNSelfTriggeringEnv(PBNEnv):
    def __init__(self, render_mode="human", render_no_cache=False, PBN_data=None, logic_func_data=None, name=None,
                 goal_config=None, reward_config=None, gamma=0.99, T=5, device=None, seed=None):
        super().__init__(render_mode, render_no_cache, PBN_data, logic_func_data, name, goal_config, reward_config,
                         device=device, seed=seed)
        self.gamma = gamma
        self.T = T  # default 5 as self_triggering.py:28 (the PBCN variant defaults to None = no cap, :107)
        # "Reward hardcode" of the reference (:50-53); PBNEnv._get_reward does not read them
        self.successful_reward, self.wrong_attractor_cost, self.action_cost = 1, 0, 1
        self.primitive_action_space = spaces.Discrete(self.PBN.N + 1)
        self.prob_space = spaces.Discrete(10, start=1)
        self.action_space = spaces.Tuple((self.primitive_action_space, self.prob_space))
        self.discrete_action_space = spaces.Discrete(self.primitive_action_space.n * self.prob_space.n)

    def _st_image(self):
        return self._image(("st", self.gamma, self.T), lambda: engine.EnvImage(
            self.network, abi.ENV_PBN_ST, attractors=[sorted(a) for a in self.all_attractors], targets=self._target_states(),
            gamma=self.gamma, max_interval=self.T))

    def step(self, action):
        if not self.action_space.contains(action):
            raise Exception(f"Invalid action {action}, not in action space.")
        control_action, prob = action
        total_reward, terminated, truncated, i = _macro_step(self, self._st_image(), [int(control_action), int(prob)])
        observation = self._last_state.astype(bool)
        return observation, total_reward, terminated, truncated, {
            "control_action": control_action, "interval": i, "observation_idx": state_to_idx(observation), "T": self.T}


class PBCNSelfTriggeringEnv(PBCNEnv):
    def __init__(self, render_mode="human", render_no_cache=False, PBN_data=None, logic_func_data=None, name=None,
                 goal_config=None, reward_config=None, gamma=0.99, T=None, device=None, seed=None, control="stac"):
        super().__init__(render_mode, render_no_cache, PBN_data, logic_func_data, name, goal_config, reward_config,
                         device=device, seed=seed, control=control)
        self.gamma = gamma
        self.T = T
        self.primitive_action_space = spaces.MultiBinary(self.PBN.M)
        self.primitive_action_space.dtype = bool
        self.prob_space = spaces.Discrete(10, start=1)
        self.action_space = spaces.Tuple((self.primitive_action_space, self.prob_space))
        self.discrete_action_space = spaces.Discrete((2**self.PBN.M) * self.prob_space.n)
        # "Reward hardcode" of the reference (self_triggering.py:134-137)
        self.successful_reward = 1
        self.wrong_attractor_cost = 1
        self.action_cost = 1
        self._invalidate_images()

    def _idx_to_macro_action(self, i):
        m = self.PBN.M
        return booleanize(i % (2**m), m).tolist(), i // (2**m) + 1

    def _st_image(self):
        return self._image(("st", self.gamma, self.T), lambda: engine.EnvImage(
            self.network, abi.ENV_PBCN_ST, attractors=[sorted(a) for a in self.all_attractors], targets=self._target_states(),
            n_control=self.PBN.M, control_write=self.PBN.control_mode == "write",
            successful_reward=self.successful_reward, wrong_attractor_cost=self.wrong_attractor_cost,
            gamma=self.gamma, max_interval=self.T))

    def step(self, action):
        if action is None:
            raise Exception("You need to provide a macro action with either `macro_action` or `macro_action_discrete`.")
        if np.isscalar(action):
            if not self.discrete_action_space.contains(action):
                raise Exception(f"Invalid action {action}, not in action space.")
            action = self._idx_to_macro_action(int(action))
        if type(action[1]) is float:
            action = (action[0], int(action[1] * 10))
        if not self.action_space.contains(action):
            raise Exception(f"Invalid action {action}, not in action space.")
        control_action, prob = action
        control = [int(bool(c)) for c in np.asarray(control_action).reshape(-1)]
        self.PBN.apply_control(control)
        total_reward, terminated, truncated, i = _macro_step(self, self._st_image(), [int(prob)] + control)
        observation = self._last_state.astype(bool)
        return observation, total_reward, terminated, truncated, {
            "control_action": control_action, "interval": i, "observation_idx": state_to_idx(observation), "T": self.T}
