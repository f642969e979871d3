"""PBNTargetMultiEnv and the BittnerMulti-N envs (reference: gym_PBN/envs/pbn_target_multi.py) — the env the
reference README calls the most interesting one: several genes may be flipped per step.

step(actions): flip every non-zero action's gene (a Python list is used as given, so duplicates cancel; a tensor is
unique()'d first — pbn_target_multi.py:120-131); the observation is captured BEFORE the first update (:133), then
updates run until that observation is an attracting state (:135-146); reward = 1000 if the observation matches the
FIRST cube of the target attractor (in_target returns at the first mismatching cube, :190-199) minus len(actions)
(:201-225); truncated when n_steps == horizon.  One CUDA launch per step.
"""
import random
from pathlib import Path

import numpy as np

from gym_PBN.b200 import abi, attractors as att_tools, engine
from gym_PBN.b200.gym_compat import spaces

from ._device import state_to_idx
from .bittner import utils
from .pbn_target import DEFAULT_MAX_INNER, _MELANOMA_7, PBNTargetEnv


class PBNTargetMultiEnv(PBNTargetEnv):
    metadata = {"render_modes": ["human", "dict", "PBN", "STG", "idx", "float", "target"]}
    _KIND = abi.ENV_MULTI

    def __init__(self, graph, goal_config, render_mode=None, render_no_cache=False, name=None, reward_config=None,
                 end_episode_on_success=False, all_attractors=None, max_inner_steps=DEFAULT_MAX_INNER, sample_pair=False):
        goal_config = dict(goal_config or {})
        for k in ("target_nodes", "target_node_values", "undesired_node_values", "intervene_on"):
            goal_config.setdefault(k, None)  # the multi env does not validate its goal config (:87-117)
        super().__init__(graph, goal_config, render_mode, render_no_cache, name, reward_config, end_episode_on_success,
                         all_attractors=all_attractors, max_inner_steps=max_inner_steps)
        self.action_space = spaces.MultiDiscrete(self.graph.N + 1)
        self.sample_pair = sample_pair  # False = the reference: always first attractor -> last attractor (:237-238)
        self.probabilities = []
        self.state_attractor_id = self.target_attractor_id = -1
        self._refresh_probabilities()

    def _refresh_probabilities(self):
        self.attractor_count = len(self._all_attractors)
        if self.attractor_count:
            self.probabilities = [1 / self.attractor_count] * self.attractor_count

    @PBNTargetEnv.all_attractors.setter
    def all_attractors(self, value):
        PBNTargetEnv.all_attractors.fset(self, value)
        self._refresh_probabilities()

    @property
    def attracting_states(self):
        """The wildcard-expanded set the reference builds (:438-454); only for small star counts."""
        out = set()
        for a in self._all_attractors:
            for c in a:
                out.update(att_tools.expand_cube(c))
        return out

    def in_target(self, observation):
        if not self.target:
            return False
        return att_tools.cube_matches(self.target[0], observation)  # only the first cube is ever consulted (:190-199)

    def step(self, actions, force=False):
        if self.target is None:
            raise ValueError("Target should have been initialized during env.reset()")
        if not isinstance(actions, list):
            actions = np.unique(np.asarray(actions.cpu() if hasattr(actions, "cpu") else actions).reshape(-1)).tolist()
        for a in actions:
            if not 0 <= int(a) <= self.graph.N:
                raise ValueError(f"Invalid action, no node at index {int(a) - 1}")
        self.sim.target_att[0] = self._target_index
        image = self._image(("multi", len(actions) or 1), lambda: engine.EnvImage(
            self.network, abi.ENV_MULTI, attractors=self._all_attractors, horizon=self.horizon,
            max_inner=self.max_inner_steps, dedup=False))
        acts = [int(a) for a in actions] if len(actions) else [-1]
        reward, terminated, truncated, inner = self._run_step(image, acts)
        self.last_inner_steps = inner
        observation = tuple(int(b) for b in self._last_obs)
        info = {"observation_idx": state_to_idx(observation), "observation_dict": observation,
                "inner_steps": inner, "inner_cap_hit": inner >= self.max_inner_steps}
        return observation, reward, terminated, truncated, info

    def reset(self, seed=None, options=None):
        if seed:
            self._seed(seed)
        if not self._all_attractors:
            raise ValueError("reset needs attractors (pbn_target_multi.py:232-238)")
        n = len(self._all_attractors)
        if n >= 2:
            self.state_attractor_id, self.target_attractor_id = np.random.choice(range(n), size=2, replace=False,
                                                                                 p=self.probabilities)
        if self.sample_pair and n >= 2:
            state_attractor = self._all_attractors[int(self.state_attractor_id)]
            target_attractor = self._all_attractors[int(self.target_attractor_id)]
        else:
            state_attractor, target_attractor = self._all_attractors[0], self._all_attractors[-1]
        state = list(random.choice(state_attractor))
        target = list(random.choice(target_attractor))
        for i in range(len(state)):
            if state[i] == "*":
                state[i] = random.randint(0, 1)
            if target[i] == "*":
                target[i] = random.randint(0, 1)
        self.graph.setState(state)
        self.n_steps = 0
        observation = self.graph.state_view(state)
        info = {"observation_idx": state_to_idx(observation), "observation_dict": observation}
        self.target = target_attractor
        self._target_index = self._index_of(target_attractor)
        return (tuple(state), tuple(target)), info

    def get_state(self):
        return np.array(self.graph.getState())

    def rework_probas_epoch(self, len_recap):
        pass

    def rework_probas(self, episode_len):
        """Curriculum hook (:159-181): make attractor pairs with long episodes more likely, short ones less."""
        eps = 1.0 / self.attractor_count
        lo, hi = 0.01 / self.attractor_count, 0.5
        s, t = int(self.state_attractor_id), int(self.target_attractor_id)
        if episode_len < 20:
            for k in (s, t):
                self.probabilities[k] = max(self.probabilities[k] - eps, lo)
        if episode_len >= 99:
            for k in (s, t):
                self.probabilities[k] = min(self.probabilities[k] + eps, hi)
        self.probabilities = [max(lo, p) for p in self.probabilities]
        total = sum(self.probabilities)
        self.probabilities = [p / total for p in self.probabilities]


class _BittnerMulti(PBNTargetMultiEnv):
    predictor_sets_path = Path(__file__).parent / "bittner" / "data"
    genedata = predictor_sets_path / "genedata.xls"
    includeIDs = sorted(_MELANOMA_7)
    N = None
    NAME = None
    _GOAL = {"target_nodes": list(_MELANOMA_7), "intervene_on": list(_MELANOMA_7),
             "target_node_values": ((1, 1, 1, 1, 1, 1, 0),), "undesired_node_values": tuple()}

    def __init__(self, render_mode="human", render_no_cache=False, name=None, horizon=100, reward_config=None,
                 end_episode_on_success=True, all_attractors=None, max_inner_steps=DEFAULT_MAX_INNER, device=None,
                 seed=None, sample_pair=False, predictor_set=None):
        graph = utils.spawn(file=self.genedata, total_genes=self.N, include_ids=self.includeIDs, bin_method="median",
                            n_predictors=3, predictor_sets_path=self.predictor_sets_path, device=device, seed=seed,
                            predictor_set=predictor_set)
        goal = dict(self._GOAL)
        goal["horizon"] = horizon
        super().__init__(graph, goal, render_mode, render_no_cache, name or self.NAME, reward_config,
                         end_episode_on_success, all_attractors=all_attractors, max_inner_steps=max_inner_steps,
                         sample_pair=sample_pair)
        if not self._all_attractors:
            self.all_attractors, self.attractor_source = att_tools.default_attractors(
                self.network, self.target_node_indices, seed=seed or 0)


def _statistical_attractors(self, resets=100, steps=1000, top=4):
    """The `top` most visited states over resets x steps forced updates (pbn_target_multi.py:465-487, there cached in
    data/attractors_{name}.pkl), sampled on the device."""
    return [a[0] for a in att_tools.statistical_attractors(self.network, resets, steps, top)]


_BittnerMulti.statistical_attractors = _statistical_attractors


class BittnerMulti7(_BittnerMulti):
    N = 7
    NAME = "Bittner-7"


class BittnerMulti10(_BittnerMulti):
    N = 10
    NAME = "BittnerMulti-10"


class BittnerMulti20(_BittnerMulti):
    N = 20
    NAME = "BittnerMulti-20"


class BittnerMulti25(_BittnerMulti):
    N = 25
    NAME = "BittnerMulti-25"


class BittnerMulti28(_BittnerMulti):
    N = 28
    NAME = "BittnerMulti-28"

    def __init__(self, render_mode="human", render_no_cache=False, name="Bittner-28", horizon=100, reward_config=None,
                 end_episode_on_success=False, **kw):
        super().__init__(render_mode, render_no_cache, name, horizon, reward_config, end_episode_on_success, **kw)


class BittnerMulti30(_BittnerMulti):
    N = 30
    NAME = "BittnerMulti-30"


class BittnerMulti50(_BittnerMulti):
    N = 50
    NAME = "BittnerMulti-50"


class BittnerMulti70(_BittnerMulti):
    N = 70
    NAME = "Bittner-70"
    _GOAL = {"target_nodes": list(_MELANOMA_7), "intervene_on": [234237], "target_node_values": ((0,) * 7,),
             "undesired_node_values": tuple()}


class BittnerMulti100(BittnerMulti70):
    N = 100
    NAME = "Bittner-100"


class BittnerMulti200(BittnerMulti70):
    N = 200
    NAME = "Bittner-200"


Bittner200 = BittnerMulti200  # the name this class carries in the reference module (pbn_target_multi.py:376)


class BittnerMultiGeneral(_BittnerMulti):
    def __init__(self, N, **kw):
        self.N = N
        self.NAME = f"BittnerMulti-{N}"
        super().__init__(**kw)
