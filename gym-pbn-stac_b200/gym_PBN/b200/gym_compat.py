"""gymnasium when it is installed; otherwise a minimal stand-in with the same surface the envs use
(Env, spaces.{Discrete, MultiBinary, MultiDiscrete, Tuple}, register, make, TimeLimit).  API plumbing only."""
import importlib

import numpy as np

try:  # pragma: no cover - depends on the image
    import gymnasium as _gym
    from gymnasium import spaces  # noqa: F401

    Env = _gym.Env
    register = _gym.register
    make = _gym.make
    HAVE_GYMNASIUM = True
except Exception:
    HAVE_GYMNASIUM = False

    class Env:
        metadata = {"render_modes": []}
        observation_space = None
        action_space = None
        render_mode = None

        def reset(self, *, seed=None, options=None):
            raise NotImplementedError

        def step(self, action):
            raise NotImplementedError

        def render(self):
            return None

        def close(self):
            pass

        @property
        def unwrapped(self):
            return self

    class _Space:
        dtype = None

        def __init__(self):
            self._rng = np.random.default_rng()

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

    class Discrete(_Space):
        def __init__(self, n, start=0):
            super().__init__()
            self.n, self.start, self.dtype = int(n), int(start), np.int64

        def contains(self, x):
            if isinstance(x, (list, tuple, dict)) or (isinstance(x, np.ndarray) and x.shape != ()):
                return False
            try:
                xi = int(x)
            except Exception:
                return False
            return xi == x and self.start <= xi < self.start + self.n

        def sample(self):
            return int(self.start + self._rng.integers(self.n))

        def __repr__(self):
            return f"Discrete({self.n}" + (f", start={self.start})" if self.start else ")")

    class MultiBinary(_Space):
        def __init__(self, n):
            super().__init__()
            self.n, self.dtype = n, np.int8
            self.shape = (n,) if np.isscalar(n) else tuple(n)

        def contains(self, x):
            a = np.asarray(x)
            return a.shape == self.shape and bool(np.all((a == 0) | (a == 1)))

        def sample(self):
            return self._rng.integers(0, 2, size=self.shape).astype(self.dtype)

        def __repr__(self):
            return f"MultiBinary({self.n})"

    class MultiDiscrete(_Space):
        def __init__(self, nvec):
            super().__init__()
            self.nvec, self.dtype = np.atleast_1d(np.asarray(nvec, dtype=np.int64)), np.int64
            self.shape = self.nvec.shape

        def contains(self, x):
            a = np.asarray(x)
            return a.shape == self.shape and bool(np.all((a >= 0) & (a < self.nvec)))

        def sample(self):
            return (self._rng.random(self.shape) * self.nvec).astype(self.dtype)

        def __repr__(self):
            return f"MultiDiscrete({self.nvec.tolist()})"

    class Tuple(_Space):
        def __init__(self, spaces_):
            super().__init__()
            self.spaces = tuple(spaces_)

        def contains(self, x):
            return (isinstance(x, (tuple, list)) and len(x) == len(self.spaces)
                    and all(s.contains(v) for s, v in zip(self.spaces, x)))

        def sample(self):
            return tuple(s.sample() for s in self.spaces)

        def __len__(self):
            return len(self.spaces)

        def __getitem__(self, i):
            return self.spaces[i]

    class _Spaces:
        pass

    spaces = _Spaces()
    spaces.Discrete, spaces.MultiBinary, spaces.MultiDiscrete, spaces.Tuple = Discrete, MultiBinary, MultiDiscrete, Tuple
    spaces.Space = _Space

    class TimeLimit(Env):
        """max_episode_steps of the registration: sets truncated once the step budget is spent."""

        def __init__(self, env, max_episode_steps):
            self.env, self._max, self._t = env, int(max_episode_steps), 0

        def __getattr__(self, name):
            return getattr(self.env, name)

        @property
        def unwrapped(self):
            return self.env

        def reset(self, **kw):
            self._t = 0
            return self.env.reset(**kw)

        def step(self, action):
            obs, r, term, trunc, info = self.env.step(action)
            self._t += 1
            return obs, r, term, trunc or self._t >= self._max, info

        def render(self, *a, **k):
            return self.env.render(*a, **k)

        def close(self):
            return self.env.close()

    registry = {}

    def register(id, entry_point=None, max_episode_steps=None, nondeterministic=False, kwargs=None, **_):
        registry[id] = dict(entry_point=entry_point, max_episode_steps=max_episode_steps,
                            nondeterministic=nondeterministic, kwargs=kwargs or {})

    def make(id, **kwargs):
        if id not in registry:
            raise KeyError(f"No registered env with id: {id}")
        spec = registry[id]
        ep = spec["entry_point"]
        if isinstance(ep, str):
            mod, attr = ep.split(":")
            ep = getattr(importlib.import_module(mod), attr)
        env = ep(**{**spec["kwargs"], **kwargs})
        if spec["max_episode_steps"]:
            env = TimeLimit(env, spec["max_episode_steps"])
        return env
