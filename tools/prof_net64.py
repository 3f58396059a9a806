"""scratch: 64-node network, specialised kernels, a few launches for ncu"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import or_gym_inventory_b200 as pkg
G = pkg.synthetic_graph(64)
N = 1 << 17
env = pkg.NetInvMgmtMasterEnv(graph=G, backlog=False, num_envs=N, device="cuda:0", specialise=True)
E = len(env.reorder_links)
a = torch.rand((N, E), device="cuda") * 100
env.reset(seed=1)
for _ in range(8): env.step(a)
torch.cuda.synchronize()
print("ok")
