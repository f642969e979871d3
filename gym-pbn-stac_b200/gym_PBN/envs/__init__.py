"""Environment classes of the drop-in package.  `from gym_PBN.envs import PBNEnv, Bittner100, ...` works as in the
reference; classes are resolved lazily from the table below so importing the package does not touch CUDA."""
import importlib

_EXPORTS = {
    "pbn_env": ["PBNEnv"],
    "pbcn_env": ["PBCNEnv"],
    "sampled_data": ["PBNSampledDataEnv", "PBCNSampledDataEnv"],
    "self_triggering": ["PBNSelfTriggeringEnv", "PBCNSelfTriggeringEnv"],
    "pbn_target": ["PBNTargetEnv"] + [f"Bittner{n}" for n in (7, 10, 28, 30, 50, 70, 100, 200)],
    "pbn_target_multi": ["PBNTargetMultiEnv", "BittnerMultiGeneral"]
    + [f"BittnerMulti{n}" for n in (7, 10, 20, 25, 28, 30, 50, 70, 100, 200)],
}
_WHERE = {name: mod for mod, names in _EXPORTS.items() for name in names}
__all__ = sorted(_WHERE)


def __getattr__(name):
    if name in _WHERE:
        value = getattr(importlib.import_module(f"{__name__}.{_WHERE[name]}"), name)
        globals()[name] = value
        return value
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


def __dir__():
    return __all__
