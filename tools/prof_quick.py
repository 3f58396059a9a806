"""scratch: small launches for ncu (one family per invocation)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import or_gym_inventory_b200 as pkg
which = sys.argv[1]
if which == "net":
    N = 1 << 20
    env = pkg.NetInvMgmtBacklogEnv(num_envs=N, device="cuda:0")
    a = torch.rand((N, 11), device="cuda") * 100
    env.reset(seed=1)
    for _ in range(12): env.step(a)
    for k in range(2): env.rollout("constant", seed=6000, episode=k)
elif which == "nv":
    N = 1 << 22
    env = pkg.NewsvendorEnv(num_envs=N, device="cuda:0")
    a = torch.rand((N, 1), device="cuda") * 100
    env.reset(seed=1)
    for _ in range(8): env.step(a)
    for k in range(2): env.rollout("classic", seed=2000, episode=k)
elif which == "inv":
    N = 1 << 24
    env = pkg.InvManagementLostSalesEnv(num_envs=N, device="cuda:0")
    a = torch.randint(0, 100, (N, 3), dtype=torch.int64, device="cuda")
    env.reset(seed=1)
    for _ in range(14): env.step(a)
    for k in range(3): env.rollout("base_stock", seed=5000, episode=k)
torch.cuda.synchronize()
print("ok")
