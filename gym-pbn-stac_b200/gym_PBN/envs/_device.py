"""Shared plumbing of the single-env drop-in classes: one device `Simulator` with B = 1, kernel launches for
`step`, host-typed return values (NumPy arrays / Python ints / bools) like the reference hands out."""
import os
import random

import numpy as np
import torch

from gym_PBN.b200 import engine


class DeviceEnvMixin:
    network = None      # engine.Network
    sim = None          # engine.Simulator with B = 1
    _images = None      # compiled env descriptions keyed by flags

    def _seed(self, seed=None):
        """Reference envs seed numpy's and Python's global RNGs (pbn_env.py:89-91); the Philox key follows."""
        np.random.seed(seed)
        random.seed(seed)
        if seed is not None:
            self.sim.reseed(int(seed))

    def _next_seed(self):
        return int.from_bytes(os.urandom(8), "little")

    def _image(self, key, build):
        if self._images is None:
            self._images = {}
        if key not in self._images:
            self._images[key] = build()
        return self._images[key]

    def _invalidate_images(self):
        self._images = {}

    def _run_step(self, image, actions):
        """One K2 launch for the single env -> (reward, terminated, truncated, inner_steps)."""
        a = torch.as_tensor(np.asarray(actions, dtype=np.int32).reshape(1, -1))
        self.sim.env_step(image, a)
        out = torch.stack([self.sim.reward, self.sim.terminated.to(torch.int32), self.sim.truncated.to(torch.int32),
                           self.sim.inner]).cpu().numpy()
        return int(out[0, 0]), bool(out[1, 0]), bool(out[2, 0]), int(out[3, 0])

    def _bits(self, planes=None):
        return self.sim.unpack(planes)[0].cpu().numpy()


def state_to_idx(state):
    """Arbitrary-precision integer of the state bits, most significant first (pbn_env.py:233-236)."""
    idx = 0
    for b in state:
        idx = (idx << 1) | int(b)
    return idx
