"""PBNTargetEnv and the Bittner-N envs (reference: gym_PBN/envs/pbn_target.py).

step(action): flip gene `action-1` (0 = no action), one asynchronous update, then keep updating until the state
matches a cube of any attractor (pbn_target.py:266-271, Bittner7.is_attracting_state :562-574); reward +20 and
terminated when the state matches a cube of the episode's target attractor, else -5; truncated when
n_steps == horizon (:303-326).  The whole step — including the unbounded inner loop, here capped by
`max_inner_steps` — is one CUDA launch on the bit-packed state.

Attractors are input data: pass `all_attractors` (CABEAN cube format, tuples over {0, 1, '*'}).  The reference gets them
from the external CABEAN binary, which is not available; without a list the Bittner envs fall back to the reference's
own sampling recipe (statistical_attractors, :538-560) run on the GPU and projected on the target genes.
"""
import random
from pathlib import Path

import numpy as np

from gym_PBN.b200 import abi, attractors as att_tools, engine
from gym_PBN.b200.gym_compat import Env, spaces

from ._device import DeviceEnvMixin, state_to_idx
from .bittner import base, utils

DEFAULT_MAX_INNER = 1 << 20


def state_equals(state1, state2):
    return all(a == b for a, b in zip(state1, state2))


class PBNTargetEnv(DeviceEnvMixin, Env):
    metadata = {"render_modes": ["human", "dict", "PBN", "STG", "idx", "float", "target"]}
    _KIND = abi.ENV_TARGET

    def __init__(self, graph, goal_config, render_mode=None, render_no_cache=False, name=None, reward_config=None,
                 end_episode_on_success=False, all_attractors=None, max_inner_steps=DEFAULT_MAX_INNER):
        self.target = None
        self.graph = graph
        self.network, self.sim = graph.network, graph.sim
        goal_config = self._check_config(
            goal_config, "goal", {"target_nodes", "target_node_values", "undesired_node_values", "intervene_on"})
        if goal_config is None:
            raise ValueError("Target nodes, target values and intervention nodes need to be specified.")
        self.target_nodes = goal_config["target_nodes"]
        self.target_node_values = goal_config.get("target_node_values")
        self.undesired_node_values = goal_config.get("undesired_node_values")
        self.intervene_on = goal_config.get("intervene_on")
        self.end_episode_on_success = end_episode_on_success
        self.horizon = goal_config["horizon"] if "horizon" in goal_config else 100

        reward_config = self._check_config(
            reward_config, "reward", {"successful_reward", "wrong_attractor_cost", "action_cost"},
            default_values={"successful_reward": 10, "wrong_attractor_cost": 2, "action_cost": 1})
        self.successful_reward = reward_config.get("successful_reward", 10)
        self.wrong_attractor_cost = reward_config.get("wrong_attractor_cost", 2)
        self.action_cost = reward_config.get("action_cost", 1)

        self.observation_space = spaces.MultiBinary(self.graph.N)
        self.action_space = spaces.Discrete(self.graph.N + 1)  # intervention nodes + no action
        self.name = name
        self.render_mode = render_mode
        self.render_no_cache = render_no_cache
        self.max_inner_steps = int(max_inner_steps)
        self._all_attractors = [list(a) for a in (all_attractors or [])]
        self._target_index = -1
        self.last_inner_steps = 0
        self.attractor_source = "given" if self._all_attractors else None

    # ---- config / attractors ---------------------------------------------------------------------------------
    def _check_config(self, config, _type, required_keys, default_values=None):
        if config:
            missing = set(required_keys) - set(config.keys())
            if len(missing) > 1:
                raise ValueError(
                    f"Invalid {_type} config provided. The following required values are missing: {', '.join(sorted(missing))}."
                )
            return config
        return default_values

    @property
    def all_attractors(self):
        return self._all_attractors

    @all_attractors.setter
    def all_attractors(self, value):
        self._all_attractors = [list(a) for a in value]
        self._invalidate_images()

    @property
    def n_steps(self):
        return int(self.sim.n_steps[0])

    @n_steps.setter
    def n_steps(self, v):
        self.sim.n_steps[0] = int(v)

    def _env_image(self, force=False):
        return self._image(("step", bool(force)), lambda: engine.EnvImage(
            self.network, self._KIND, attractors=self._all_attractors, horizon=self.horizon,
            max_inner=self.max_inner_steps, force=force))

    @property
    def env_image(self):
        return self._env_image(False) if self._all_attractors else None

    @property
    def target_node_indices(self):
        ids = self.graph.getIDs()
        return [ids.index(t) for t in self.target_nodes]

    def is_attracting_state(self, state):
        state = tuple(state)
        if not self._all_attractors:
            return True
        return any(att_tools.cube_matches(c, state) for a in self._all_attractors for c in a)

    def in_target(self, observation):
        if self.target is None:
            raise ValueError("Target should have been initialized during env.reset()")
        return any(att_tools.cube_matches(c, observation) for c in self.target)

    def setTarget(self, target):
        self.target = target
        self._target_index = self._index_of(target)

    # ---- host-side helpers of the reference class (the device kernel computes the same reward inside env.step) -----
    def _to_map(self, state):
        """pbn_target.py:282-287: positional state -> {gene ID: value}."""
        getIDs = getattr(self.graph, "getIDs", None)
        if getIDs is not None and type(state) is not dict:
            state = dict(zip(getIDs(), state))
        return state

    def _get_reward(self, observation, action):
        """(reward, terminated, truncated) of an observation, pbn_target.py:303-326: +20 and terminated inside the target
        attractor, -5 otherwise; truncated at the horizon.  `observation` may be a mapping keyed by gene ID or a sequence."""
        values = tuple(observation.values()) if hasattr(observation, "values") else tuple(observation)
        hit = self.in_target(values)
        return (20, True, self.n_steps == self.horizon) if hit else (-5, False, self.n_steps == self.horizon)

    def compute_attractors(self):
        """Attractors of the asynchronous dynamics as a list of sets of state tuples (pbn_target.py:393-408, there through
        an explicit networkx STG; here the exhaustive device search, networks up to 32 genes)."""
        return att_tools.attractor_state_sets(self.network, list_limit=1 << 22)

    def _nx_attractors_to_tuples(self, attractors):
        """networkx attracting components over "[0 1 ...]" labels or state tuples -> list of sets of int tuples."""
        out = []
        for attractor in attractors:
            states = set()
            for state in attractor:
                if isinstance(state, str):
                    state = state.lstrip("[").rstrip("]").split()
                states.add(tuple(int(x) for x in state))
            out.append(states)
        return out

    def attractor_checker(self, stg, state, depth):
        """Depth-limited expansion of the state-transition graph around `state` (pbn_target.py:111-122)."""
        if depth < 1:
            return stg
        stg.add_node(state)
        next_states = self.graph.getNextStates(state)
        stg.add_nodes_from(next_states.keys())
        for ns in next_states:
            stg.add_edge(state, ns)
            stg = self.attractor_checker(stg, ns, depth - 1)
        return stg

    def dep_is_attracting_state(self, state):
        return True  # the reference's on-line heuristic is disabled by this very line (pbn_target.py:138-139)

    def also_dep_is_attracting_state(self, state):
        raise ValueError("You are not supposed to be using me")  # pbn_target.py:124-125

    def _index_of(self, attractor):
        for i, a in enumerate(self._all_attractors):
            if a is attractor or list(a) == list(attractor):
                return i
        raise ValueError("the target must be one of env.all_attractors")

    # ---- gym API ---------------------------------------------------------------------------------------------
    def step(self, action=0, force=False):
        if self.target is None:
            raise ValueError("Target should have been initialized during env.reset()")
        action = int(action)
        if not 0 <= action <= self.graph.N:
            raise ValueError(f"Invalid action, no node at index {action - 1}")
        self.sim.target_att[0] = self._target_index
        reward, terminated, truncated, inner = self._run_step(self._env_image(force), [action])
        self.last_inner_steps = inner
        observation = self.graph.state_view(self._last_state)
        info = {"observation_idx": state_to_idx(observation), "observation_dict": observation,
                "inner_steps": inner, "inner_cap_hit": inner >= self.max_inner_steps}
        return self._last_state.astype(np.int64), reward, terminated, truncated, info

    def reset(self, seed=None, options=None):
        if seed:
            self._seed(seed)
        if len(self._all_attractors) < 2:
            raise ValueError("reset needs at least two attractors (random.sample(all_attractors, 2), pbn_target.py:333)")
        state_attractor, target_attractor = random.sample(self._all_attractors, 2)
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
        return np.array(list(self.graph.getState().values()))

    def render(self, mode=None):
        mode = self.render_mode if not mode else mode
        if mode == "human":
            return self.get_state()
        if mode == "dict":
            return self.graph.getState()
        if mode == "PBN":
            return self.graph.printGraph()
        if mode == "STG":
            return self.graph.genSTG()  # explicit graph: small networks only (bittner/base.py genSTG)
        if mode == "idx":
            return state_to_idx(self.graph.getState())
        if mode == "float":
            return [float(x) for x in self.graph.getState()]
        if mode == "target":
            state = self.graph.getState()
            return [state[node] for node in self.target_nodes]
        if mode == "target_idx":
            return state_to_idx(self.render(mode="target"))

    def getTargetIdx(self):
        state = self.graph.getState()
        return state_to_idx([state[node] for node in self.target_nodes])

    def _state_to_idx(self, state):
        if isinstance(state, dict):
            state = list(state.values())
        return state_to_idx(state)

    def _seed(self, seed=None):
        super()._seed(seed)

    def close(self):
        self._invalidate_images()


# ------------------------------------------------------------------------------------------------------ Bittner family
_MELANOMA_7 = [234237, 324901, 759948, 25485, 266361, 108208, 130057]


class _BittnerTarget(PBNTargetEnv):
    predictor_sets_path = Path(__file__).parent / "bittner" / "data"
    genedata = predictor_sets_path / "genedata.xls"
    includeIDs = list(_MELANOMA_7)
    N = None
    NAME = None
    _GOAL = None

    def _build(self, render_mode, render_no_cache, name, horizon, reward_config, end_episode_on_success, all_attractors,
               max_inner_steps, device, seed, predictor_set=None):
        # predictor_set: None = shipped set of this size if upstream ships one, else fit `{N}_3_median` on the GPU;
        # "fit" = always the reference's route (bittner/utils.py:54-80); or the name of a shipped set
        graph = utils.spawn(file=self.genedata, total_genes=self.N, include_ids=self.includeIDs, bin_method="median",
                            n_predictors=3, predictor_sets_path=self.predictor_sets_path, device=device, seed=seed,
                            predictor_set=predictor_set)
        goal = dict(self._GOAL)
        goal["horizon"] = horizon
        PBNTargetEnv.__init__(self, graph, goal, render_mode, render_no_cache, name or self.NAME, reward_config,
                              end_episode_on_success, all_attractors=all_attractors, max_inner_steps=max_inner_steps)
        if not self._all_attractors:
            # no CABEAN here: exact terminal SCCs of the asynchronous STG (N <= 28); beyond that sampled + verified closed cube
            # sets (b200/attractors.py: verified_attractors); the reference's own sampling recipe projected on the target
            # genes only when that route finds fewer than two attractors (attractor_source says which)
            self.all_attractors, self.attractor_source = att_tools.default_attractors(
                self.network, self.target_node_indices, seed=seed or 0)


class Bittner70(_BittnerTarget):
    N = 70
    NAME = "Bittner-70"
    _GOAL = {"target_nodes": list(_MELANOMA_7), "intervene_on": [234237], "target_node_values": ((0,) * 7,),
             "undesired_node_values": tuple()}

    def __init__(self, render_mode="human", render_no_cache=False, name=None, horizon=69, reward_config=None,
                 end_episode_on_success=True, all_attractors=None, max_inner_steps=DEFAULT_MAX_INNER, device=None, seed=None,
                 predictor_set=None):
        self._build(render_mode, render_no_cache, name, horizon, reward_config, end_episode_on_success, all_attractors,
                    max_inner_steps, device, seed, predictor_set)


class Bittner100(Bittner70):
    N = 100
    NAME = "Bittner-100"


class Bittner200(Bittner70):
    N = 200
    NAME = "Bittner-200"


class Bittner7(_BittnerTarget):
    includeIDs = sorted(_MELANOMA_7)
    N = 7
    NAME = "Bittner-7"
    _GOAL = {"target_nodes": list(_MELANOMA_7), "intervene_on": list(_MELANOMA_7),
             "target_node_values": ((1, 1, 1, 1, 1, 1, 0),), "undesired_node_values": tuple()}

    def __init__(self, render_mode="human", render_no_cache=False, name=None, horizon=100, reward_config=None,
                 end_episode_on_success=True, all_attractors=None, max_inner_steps=DEFAULT_MAX_INNER, device=None, seed=None,
                 predictor_set=None):
        self._build(render_mode, render_no_cache, name, horizon, reward_config, end_episode_on_success, all_attractors,
                    max_inner_steps, device, seed, predictor_set)
        self.target_nodes = sorted(_MELANOMA_7)  # pbn_target.py:534 (the seven melanoma genes, sorted)
        self.target_node_values = self.all_attractors[-1]
        self.target_attractor = len(self.all_attractors) - 1  # last one (pbn_target.py:534-536)

    def statistical_attractors(self, resets=100, steps=1000, top=4):
        return [a[0] for a in att_tools.statistical_attractors(self.network, resets, steps, top)]


class Bittner10(Bittner7):
    N = 10
    NAME = "Bittner-10"


class Bittner30(Bittner7):
    N = 30
    NAME = "Bittner-30"


class Bittner50(Bittner7):
    N = 50
    NAME = "Bittner-50"


class Bittner28(Bittner7):
    N = 28
    NAME = "Bittner-28"

    def __init__(self, render_mode="human", render_no_cache=False, name="Bittner-28", horizon=100, reward_config=None,
                 end_episode_on_success=False, all_attractors=None, max_inner_steps=DEFAULT_MAX_INNER, device=None, seed=None,
                 predictor_set=None):
        super().__init__(render_mode, render_no_cache, name, horizon, reward_config, end_episode_on_success,
                         all_attractors, max_inner_steps, device, seed, predictor_set)
