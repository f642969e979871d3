"""gym_PBN — B200-native drop-in for jakub-zarzycki2022/gym-PBN-stac: same import name, same env ids
(reference registrations: gym_PBN/__init__.py:3-134), CUDA kernels underneath (gym_PBN.b200).

    import gym_PBN; env = gym_PBN.make("gym-PBN/Bittner-100-v0")         # or gymnasium.make when it is installed
    vec = gym_PBN.make_vec("gym-PBN/Bittner-100-v0", num_envs=65536)     # batched, torch tensors on device
"""
from gym_PBN.b200.gym_compat import HAVE_GYMNASIUM, make, register

__version__ = "0.1.0"

_E = "gym_PBN.envs"
register(id="gym-PBN/PBN-v0", entry_point=f"{_E}:PBNEnv")
register(id="gym-PBN/PBN-target-v0", entry_point=f"{_E}:PBNTargetEnv")
# named in the reference README (:19) and in north_star but never registered there
register(id="gym-PBN/PBN-target_multi-v0", entry_point=f"{_E}:PBNTargetMultiEnv")
register(id="gym-PBN/PBN-sampled-data-v0", entry_point=f"{_E}:PBNSampledDataEnv")
register(id="gym-PBN/PBN-self-triggering-v0", entry_point=f"{_E}:PBNSelfTriggeringEnv")
register(id="gym-PBN/PBCN-v0", entry_point=f"{_E}:PBCNEnv")
register(id="gym-PBN/PBCN-sampled-data-v0", entry_point=f"{_E}:PBCNSampledDataEnv")
register(id="gym-PBN/PBCN-self-triggering-v0", entry_point=f"{_E}:PBCNSelfTriggeringEnv")
for _n in (7, 10, 28, 30, 50, 70, 100, 200):
    # the reference registers 7/28/30/70 (:7-35); Bittner100/200 classes exist there (pbn_target.py:464-471) and
    # `Bittner-200-v0` is what example.py:53 asks for, so the whole family is registered here
    register(id=f"gym-PBN/Bittner-{_n}-v0", entry_point=f"{_E}:Bittner{_n}", nondeterministic=True, max_episode_steps=100)
for _n, _cls in ((7, 7), (10, 10), (20, 20), (25, 25), (28, 28), (30, 28), (50, 50)):  # Multi-30 -> BittnerMulti28 as in :115-120
    register(id=f"gym-PBN/BittnerMulti-{_n}-v0", entry_point=f"{_E}:BittnerMulti{_cls}", nondeterministic=True,
             max_episode_steps=100)
register(id="gym-PBN/BittnerMultiGeneral-v0", entry_point=f"{_E}:BittnerMultiGeneral", nondeterministic=True,
         max_episode_steps=100)


def make_vec(id, num_envs, **kwargs):
    """Batched counterpart of make(): a PBNVectorEnv over `num_envs` copies, tensors stay on the device."""
    from gym_PBN.b200.vector_env import make_vec as _mv

    return _mv(id, num_envs, **kwargs)
