"""Drop-in for the reference's `network_management_custom.py`: the same env with the 3-retailer default
graph (network_management_custom.py:113-139)."""
from .network_management import (NetInvMgmtBacklogEnv as _B, NetInvMgmtLostSalesEnv as _L,
                                 NetInvMgmtMasterEnv as _M)


class NetInvMgmtMasterEnv(_M):
    _default_kind = "custom"


class NetInvMgmtBacklogEnv(_B):
    _default_kind = "custom"


class NetInvMgmtLostSalesEnv(_L):
    _default_kind = "custom"
