"""PBCNEnv — PBN with control nodes (reference: gym_PBN/envs/pbcn_env.py).

The action space is MultiBinary(M) but `step` flips node `action` exactly like PBNEnv (pbcn_env.py:67-80); the reward
is `successful_reward` on a target state, else -wrong_attractor_cost x (#attractors containing the state)
(pbcn_env.py:52-65).  One CUDA launch per step.
"""
import numpy as np

from gym_PBN.b200 import abi
from gym_PBN.b200.gym_compat import spaces

from ._device import state_to_idx
from .common.pbcn import PBCN
from .pbn_env import PBNEnv


class PBCNEnv(PBNEnv):
    metadata = {"render_modes": ["human", "PBN", "STG", "funcs", "idx", "float"]}
    _CORE = PBCN
    _KIND = abi.ENV_PBCN

    def __init__(self, render_mode="human", render_no_cache=False, PBN_data=None, logic_func_data=None, name=None,
                 goal_config=None, reward_config=None, device=None, seed=None, control="stac"):
        super().__init__(render_mode, render_no_cache, PBN_data, logic_func_data, name, goal_config, reward_config,
                         device=device, seed=seed, control=control)
        self.observation_space = spaces.MultiBinary(self.PBN.N)
        self.observation_space.dtype = bool
        self.action_space = spaces.MultiBinary(self.PBN.M)
        self.action_space.dtype = bool
        self.discrete_action_space = spaces.Discrete(2**self.PBN.M)
        self.target_nodes = set(self._raw_target)  # pbcn_env.py:46 keeps the raw target, without the attractor union

    def getTargetIdx(self):
        return int(tuple(int(v) for v in self.PBN.state) in self.target_nodes)

    def _get_reward(self, observation):
        """(reward, terminated, truncated) of an observation, pbcn_env.py:52-65 — the one reward that reads reward_config."""
        observation_tuple = tuple(int(v) for v in observation)
        if observation_tuple in self.target_nodes:
            return self.successful_reward, True, False
        matched = sum(observation_tuple in attractor for attractor in self.all_attractors)
        return -self.wrong_attractor_cost * matched, False, False

    @staticmethod
    def _flip_index(action):
        a = np.asarray(action).reshape(-1)
        if a.size != 1:
            raise Exception(f"Invalid action {action}: PBCNEnv.step flips a single node index")
        return int(a[0])

    def step(self, action=0):
        a = self._flip_index(action)
        if not 0 <= a < self.PBN.N:
            raise Exception(f"Invalid action {action}, not in action space.")
        reward, terminated, truncated, _ = self._run_step(self._env_image(), [a])
        observation = self._last_state.astype(bool)
        return observation, reward, terminated, truncated, {"observation_idx": state_to_idx(observation)}
