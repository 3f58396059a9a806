import os, sys, time
sys.path.insert(0, "/root/repo")
import torch
import or_gym_inventory_b200 as pkg
G = pkg.synthetic_graph(64)
N = int(os.environ.get('N64', 1 << 17))
t0 = time.time()
env = pkg.NetInvMgmtMasterEnv(graph=G, backlog=False, num_envs=N, device="cuda:0", info_level=int(os.environ.get("INFO", 1)))
print("create s", round(time.time() - t0, 1), "E", len(env.reorder_links), "J", len(env.main_nodes), "M", len(env.retail_links), "obs", env.obs_dim)
E = len(env.reorder_links)
a = torch.rand((N, E), device="cuda") * 100
env.reset(seed=1)
def timeit(fn, K=10, W=3):
    for _ in range(W): fn(0)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for k in range(K): fn(k)
    ev[1].record(); torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / K
ms = timeit(lambda k: env.step(a))
print("net64 step", f"{ms:.3f} ms  {N/ms/1e6:.4f} Gsteps/s")
ms = timeit(lambda k: env.rollout("constant", seed=6000, episode=k, order_fraction=0.02), K=3)
print("net64 rollout const", f"{ms:.3f} ms  {N*30/ms/1e6:.3f} Gsteps/s")
