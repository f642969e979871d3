"""PBNEnv — truth-table PBN with single-node flip actions (reference: gym_PBN/envs/pbn_env.py).

step(action): flip node `action` (0 = no action; the index is `action` itself, so node 0 cannot be flipped —
pbn_env.py:141-142), one asynchronous update, reward +20 and terminated when the state is a target state,
else -4 and a further -1 for a non-zero action (pbn_env.py:156-188).  All of it is one CUDA launch.
"""
import random

import numpy as np

from gym_PBN.b200 import abi, engine
from gym_PBN.b200.gym_compat import Env, spaces

from ._device import DeviceEnvMixin, state_to_idx
from .common.pbn import PBN


class PBNEnv(DeviceEnvMixin, Env):
    metadata = {"render_modes": ["human", "PBN", "STG", "funcs", "idx", "float"]}
    _CORE = PBN
    _KIND = abi.ENV_PBN
    MAX_EXHAUSTIVE_NODES = 26  # the exhaustive STG lives on the device: 2^26 states x 4 B of change masks

    def __init__(self, render_mode="human", render_no_cache=False, PBN_data=None, logic_func_data=None, name=None,
                 goal_config=None, reward_config=None, device=None, seed=None, **core_kwargs):
        self.PBN = self._CORE(PBN_data if PBN_data is not None else [], logic_func_data, device=device, seed=seed,
                              **core_kwargs)
        self.network, self.sim = self.PBN.network, self.PBN.sim

        goal_config = self._check_config(goal_config, "goal", {"target", "all_attractors"})
        if self.PBN.N <= self.MAX_EXHAUSTIVE_NODES:
            self.all_attractors = self.compute_attractors()  # the reference always recomputes them (pbn_env.py:54)
        elif goal_config and "all_attractors" in goal_config:
            self.all_attractors = [set(tuple(int(v) for v in s) for s in a) for a in goal_config["all_attractors"]]
        else:
            raise ValueError(
                f"N={self.PBN.N}: the exhaustive state-transition graph is O(2^N); pass goal_config['all_attractors']"
            )
        if goal_config is None:
            target = set(self.all_attractors[-1])
        elif "target_nodes" in goal_config:
            assert type(goal_config["target_nodes"]) is set, "Did you put multiple attractors as the target by mistake?"
            target = set(goal_config["target_nodes"])
        else:  # README-style config: {"target": {...}, "all_attractors": [...]}
            target = set(goal_config["target"])
        self._raw_target = set(tuple(int(v) for v in t) for t in target)
        self.target_nodes = set(self._raw_target)
        for attractor in self.all_attractors:
            if self.target_nodes & attractor:
                self.target_nodes = self.target_nodes.union(attractor)
        self.attracting_states = set.union(*self.all_attractors) if self.all_attractors else set()

        reward_config = self._check_config(
            reward_config, "reward", {"successful_reward", "wrong_attractor_cost", "action_cost"},
            default_values={"successful_reward": 10, "wrong_attractor_cost": 2, "action_cost": 1})
        self.successful_reward = reward_config.get("successful_reward", 10)
        self.wrong_attractor_cost = reward_config.get("wrong_attractor_cost", 2)
        self.action_cost = reward_config.get("action_cost", 1)

        self.observation_space = spaces.MultiBinary(self.PBN.N)
        self.observation_space.dtype = bool
        self.action_space = spaces.Discrete(self.PBN.N)
        self.name = name
        self.render_mode = render_mode
        self.render_no_cache = render_no_cache
        self.step_no = 0

    # ---- configuration -------------------------------------------------------------------------------------
    def _check_config(self, config, _type, required_keys, default_values=None):
        """Same leniency as the reference (pbn_env.py:93-123): raises only when MORE than one required key is missing."""
        if config:
            missing = set(required_keys) - set(config.keys())
            if len(missing) > 1:
                raise ValueError(
                    f"Invalid {_type} config provided. The following required values are missing: {', '.join(sorted(missing))}."
                )
            return config
        return default_values

    def _target_states(self):
        return sorted(self.target_nodes)

    def _env_image(self):
        return self._image("step", lambda: engine.EnvImage(
            self.network, self._KIND, attractors=[sorted(a) for a in self.all_attractors], targets=self._target_states(),
            n_control=getattr(self.PBN, "M", 0), control_write=getattr(self.PBN, "control_mode", "stac") == "write",
            successful_reward=self.successful_reward, wrong_attractor_cost=self.wrong_attractor_cost))

    @property
    def env_image(self):
        return self._env_image()

    # ---- gym API -------------------------------------------------------------------------------------------
    def is_attracting_state(self, state):
        return True  # pbn_env.py:19-21

    def step(self, action):
        if not self.action_space.contains(action):
            raise Exception(f"Invalid action {action}, not in action space.")
        reward, terminated, truncated, _ = self._run_step(self._env_image(), [int(action)])
        observation = self._last_state.astype(bool)
        return observation, reward, terminated, truncated, {"observation_idx": state_to_idx(observation)}

    def reset(self, seed=None, options=None):
        if seed is not None:
            self._seed(seed)
        if options is not None and "state" in options:
            state = options["state"]
        else:
            state = random.choice(tuple(self.attracting_states))
        attr = None
        while attr is None or len(attr) > 10:  # the first pick is overwritten, as in pbn_env.py:195-206
            attr = random.choice(self.all_attractors)
        state = random.choice(tuple(attr))
        observation = self.PBN.reset(state)
        if tuple(int(v) for v in observation) not in self.attracting_states:
            raise ValueError("state initial state should be an attractor")
        self.step_no = 0
        return observation, {"observation_idx": state_to_idx(observation)}

    def set(self, new_state):
        self.PBN.state = np.array(new_state)

    def _get_reward(self, observation, action):
        """(reward, terminated, truncated) of an observation, pbn_env.py:156-188: +20 and terminated in a target state,
        else -4 (the state must be attracting, ValueError otherwise) and -1 more for a non-null action."""
        observation_tuple = tuple(int(v) for v in observation)
        if observation_tuple in self.target_nodes:
            return 20, True, False
        if not self.is_attracting_state(observation):
            raise ValueError
        return -4 - (1 if action != 0 else 0), False, False

    def _nx_attractors_to_tuples(self, attractors):
        return [set(tuple(int(x) for x in state.lstrip("[").rstrip("]").split()) for state in attractor) for attractor in attractors]

    def clip(self, gene_i):
        self.PBN.clip(gene_i)

    def render(self, mode=None):
        mode = self.render_mode if mode is None else mode
        if mode == "human":
            return self.PBN.state
        if mode == "PBN":
            return self.PBN.print_PBN(False)
        if mode == "STG":
            return self.PBN.print_STG(False)
        if mode == "funcs":
            return self.PBN.print_functions()
        if mode == "idx":
            return state_to_idx(self.PBN.state)
        if mode == "float":
            return [float(x) for x in self.PBN.state]

    def _state_to_idx(self, state):
        return state_to_idx(np.asarray(state).astype(int).tolist())

    def compute_attractors(self):
        return self.PBN.attractors(self.MAX_EXHAUSTIVE_NODES)

    def close(self):
        self._invalidate_images()
