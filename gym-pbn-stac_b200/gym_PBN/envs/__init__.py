from gym_PBN.envs.pbcn_env import PBCNEnv
from gym_PBN.envs.pbn_env import PBNEnv
from gym_PBN.envs.pbn_target import (
    Bittner7,
    Bittner10,
    Bittner28,
    Bittner30,
    Bittner50,
    Bittner70,
    Bittner100,
    Bittner200,
    PBNTargetEnv,
)
from gym_PBN.envs.pbn_target_multi import (
    BittnerMulti7,
    BittnerMulti10,
    BittnerMulti20,
    BittnerMulti25,
    BittnerMulti28,
    BittnerMulti30,
    BittnerMulti50,
    BittnerMulti70,
    BittnerMulti100,
    BittnerMulti200,
    BittnerMultiGeneral,
    PBNTargetMultiEnv,
)
from gym_PBN.envs.sampled_data import PBCNSampledDataEnv, PBNSampledDataEnv
from gym_PBN.envs.self_triggering import PBCNSelfTriggeringEnv, PBNSelfTriggeringEnv
