"""Shared plumbing of the single-env drop-in classes: one device `Simulator` with B = 1, kernel launches for
`step`, host-typed return values (NumPy arrays / Python ints / bools) like the reference hands out."""
import ctypes as C
import os
import random

import numpy as np
import torch

from gym_PBN.b200 import abi, engine


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

    def replay_draws(self, ints, dbls):
        """Queue recorded draws of the reference's RNGs for the NEXT kernel launch of this env: ints =
        results of randint / choice, dbls = uniforms, each in the reference's call order.  One entry per launch; launches
        without a queued entry draw from the env's Philox stream."""
        if getattr(self, "_replays", None) is None:
            self._replays = []
        self._replays.append(engine.Replay(ints, dbls, self.sim.device, B=1))

    def _single_io(self, K):
        """Pinned host staging for the single env: actions up, {reward, flags, inner, state, obs} down in ONE read-back."""
        ios = self.__dict__.setdefault("_ios", {})  # one staging block per action width: built once each, never rebuilt
        io = ios.get(K)
        if io is None:
            sim, w32 = self.sim, self.network.w32
            host = torch.empty(12 + 8 * w32, dtype=torch.uint8).pin_memory()
            acts_host = torch.empty(K, dtype=torch.int32).pin_memory()
            io = {"K": K, "host": host, "np": host.numpy(), "acts_host": acts_host, "acts_np": acts_host.numpy(),
                  "acts_dev": torch.empty((1, K), dtype=torch.int32, device=sim.device),
                  "scratch": torch.empty(3 + 2 * w32, dtype=torch.int32, device=sim.device)}
            p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
            io["step_args"] = lambda image, d, stream: (
                image.handle, p(sim.state), p(sim.n_steps), p(sim.target_att), p(io["acts_dev"]), K, p(sim.obs_state),
                p(sim.reward), p(sim.terminated), p(sim.truncated), p(sim.inner), 1, sim.env0, C.byref(d), stream)
            io["fetch_args"] = (p(sim.reward), p(sim.terminated), p(sim.truncated), p(sim.inner), p(sim.state),
                                p(sim.obs_state), w32, 1, 0, p(io["scratch"]), p(host))
            io["up_args"] = (p(io["acts_dev"]), p(acts_host), 4 * K)
            ios[K] = io
        return io

    def _run_step(self, image, actions):
        """One K2 launch for the single env and one read-back -> (reward, terminated, truncated, inner_steps);
        the new state / observation bits are left in self._last_state / self._last_obs (uint8 [N])."""
        acts = np.asarray(actions, dtype=np.int32).reshape(-1)
        io = self._single_io(len(acts))
        io["acts_np"][:] = acts
        sim, lib = self.sim, abi.lib()
        queue = getattr(self, "_replays", None)
        replay = queue.pop(0) if queue else None
        d = sim._draws(replay)
        with engine.on_device(sim.device):  # the caller's current device is restored (a process may hold several GPUs)
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            abi.check(lib.pbn_upload(*io["up_args"], stream))
            abi.check(lib.pbn_env_step(*io["step_args"](image, d, stream)))
            abi.check(lib.pbn_fetch_step_host(*io["fetch_args"], stream))
        sim.launches += 2
        buf, n, w4 = io["np"], self.network.n, 4 * self.network.w32
        head = buf[:12].view(np.int32)
        self._last_state = np.unpackbits(buf[12:12 + w4], bitorder="little")[:n]
        self._last_obs = np.unpackbits(buf[12 + w4:12 + 2 * w4], bitorder="little")[:n]
        return int(head[0]), bool(head[2] & 0xFF), bool((head[2] >> 8) & 0xFF), int(head[1])

    def _bits(self, planes=None):
        return self.sim.unpack(planes)[0].cpu().numpy()


def state_to_idx(state):
    """Arbitrary-precision integer of the state bits, most significant first (pbn_env.py:233-236)."""
    idx = 0
    for b in state:
        idx = (idx << 1) | int(b)
    return idx
