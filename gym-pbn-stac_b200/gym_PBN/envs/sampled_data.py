"""Temporally extended actions (reference: gym_PBN/envs/sampled_data.py): an action is (primitive, interval); the
primitive is applied and one update made `interval` times, rewards are summed inside the kernel — so envs with
different intervals (config 5) simply run different trip counts in the same launch.
"""
import numpy as np

from gym_PBN.b200 import abi
from gym_PBN.b200.gym_compat import spaces
from gym_PBN.utils import booleanize

from ._device import state_to_idx
from .pbcn_env import PBCNEnv
from .pbn_env import PBNEnv


class PBNSampledDataEnv(PBNEnv):
    _KIND = abi.ENV_PBN_SD

    def __init__(self, render_mode="human", render_no_cache=False, PBN_data=None, logic_func_data=None, name=None,
                 goal_config=None, reward_config=None, gamma=0.99, T=None, device=None, seed=None):
        super().__init__(render_mode, render_no_cache, PBN_data, logic_func_data, name, goal_config, reward_config,
                         device=device, seed=seed)
        self.gamma = gamma
        self.T = T if T is not None else 2**self.PBN.N
        self.primitive_action_space = spaces.Discrete(self.PBN.N + 1)
        self.interval_space = spaces.Discrete(self.T, start=1)
        self.action_space = spaces.Tuple((self.primitive_action_space, self.interval_space))
        self.discrete_action_space = spaces.Discrete(self.primitive_action_space.n * self.interval_space.n)

    def step(self, action):
        if not self.action_space.contains(action):
            raise Exception(f"Invalid action {action}, not in action space.")
        control_action, interval = int(action[0]), int(action[1])
        reward, terminated, truncated, _ = self._run_step(self._env_image(), [control_action, interval])
        observation = self._last_state.astype(bool)
        info = {"control_action": control_action, "interval": interval - 1,  # the reference reports the last loop index
                "observation_idx": state_to_idx(observation)}
        return observation, reward, terminated, truncated, info


class PBCNSampledDataEnv(PBCNEnv):
    _KIND = abi.ENV_PBCN_SD

    def __init__(self, render_mode="human", render_no_cache=False, PBN_data=None, logic_func_data=None, name=None,
                 goal_config=None, reward_config=None, gamma=0.99, T=None, device=None, seed=None, control="stac"):
        super().__init__(render_mode, render_no_cache, PBN_data, logic_func_data, name, goal_config, reward_config,
                         device=device, seed=seed, control=control)
        self.gamma = gamma
        if T is None and self.PBN.N > 62:
            raise ValueError("T defaults to 2**N (sampled_data.py:121); give an explicit T for large networks")
        self.T = T if T is not None else 2**self.PBN.N
        self.primitive_action_space = spaces.MultiBinary(self.PBN.M)
        self.primitive_action_space.dtype = bool
        self.interval_space = spaces.Discrete(self.T, start=1)
        self.action_space = spaces.Tuple((self.primitive_action_space, self.interval_space))
        self.discrete_action_space = spaces.Discrete((2**self.PBN.M) * self.T)

    def _idx_to_macro_action(self, i):
        m = self.PBN.M
        return booleanize(i % (2**m), m).tolist(), i // (2**m) + 1

    def step(self, action):
        if action is None:
            raise Exception("You need to provide a macro action with either `macro_action` or `macro_action_discrete`.")
        if np.isscalar(action) or (isinstance(action, np.ndarray) and action.shape == ()):
            if not self.discrete_action_space.contains(action):
                raise Exception(f"Invalid action {action}, not in action space.")
            action = self._idx_to_macro_action(int(action))
        if not self.action_space.contains(action):
            raise Exception(f"Invalid action {action}, not in action space.")
        control_action, interval = action
        control = [int(bool(c)) for c in np.asarray(control_action).reshape(-1)]
        self.PBN.apply_control(control)
        reward, terminated, truncated, _ = self._run_step(self._env_image(), [int(interval)] + control)
        observation = self._last_state.astype(bool)
        info = {"control_action": control_action, "interval": int(interval), "observation_idx": state_to_idx(observation)}
        return observation, reward, terminated, truncated, info
