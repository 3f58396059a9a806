"""B200-native batched simulator for the step dynamics of jacklu2016/or-gym-inventory.

Drop-in (vectorised) replacements for the reference's three env families, same class names and
constructor keywords, each running `num_envs` instances on one GPU behind a C-ABI CUDA library:

    from or_gym_inventory_b200 import InvManagementBacklogEnv
    env = InvManagementBacklogEnv(num_envs=1 << 20, device="cuda:0")
    obs, info = env.reset(seed=0)
    obs, reward, terminated, truncated, info = env.step(actions)      # torch CUDA tensors, zero-copy
    out = env.rollout("base_stock")                                   # fused 30-period rollout

Importing the package does not load the CUDA library; constructing an env does, and raises if the
library has not been built (there is no CPU fallback).
"""
from . import _capi  # noqa: F401
from .inventory_management import (InvManagementBacklogEnv, InvManagementLostSalesEnv,  # noqa: F401
                                   InvManagementMasterEnv, InvManagementParams)

from .newsvendor import NewsvendorEnv, NewsvendorParams  # noqa: F401
from .network_management import (NetInvMgmtBacklogEnv, NetInvMgmtLostSalesEnv, NetInvMgmtMasterEnv,  # noqa: F401
                                 NetInvMgmtParams, default_graph, graph_from_spec, synthetic_graph)
from . import network_management_custom  # noqa: F401
from .adapters import SB3VecEnvAdapter, rllib_env_creator  # noqa: F401
from .sampling import sample_demand, sample_poisson_mu  # noqa: F401
from .sharding import allreduce_summary, describe_summary, shard_range  # noqa: F401
from .metrics import (REPORT_FIELDS, evaluation_report, evaluation_report_device, kth_smallest,  # noqa: F401
                      report_to_dict)

__version__ = "0.1.0"
