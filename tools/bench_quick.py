"""scratch: quick timing of all kernels (superseded by bench.py)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import or_gym_inventory_b200 as pkg

def timeit(fn, K=10, W=3):
    for _ in range(W): fn(0)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for k in range(K): fn(k)
    ev[1].record(); torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / K

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "inv"):
    N = 1 << 24
    env = pkg.InvManagementLostSalesEnv(num_envs=N, device="cuda:0")
    T = env.num_periods
    for pol in ("base_stock", "random"):
        ms = timeit(lambda k: env.rollout(pol, seed=5000, episode=k))
        print("inv rollout", pol, f"{ms:.3f} ms  {N*T/ms/1e6:.1f} Gsteps/s")
    a = torch.randint(0, 100, (N, 3), dtype=torch.int64, device="cuda")
    env.reset(seed=1)
    ms = timeit(lambda k: env.step(a), K=20)
    print("inv step", f"{ms:.3f} ms  {N/ms/1e6:.2f} Gsteps/s  {N*466/ms/1e6:.0f} GB/s alg")
    env.close(); del env, a
if which in ("all", "nv"):
    N = 1 << 24
    env = pkg.NewsvendorEnv(num_envs=N, device="cuda:0")
    T = env.step_limit
    for pol in ("classic", "order_up_to", "sS"):
        ms = timeit(lambda k: env.rollout(pol, seed=2000, episode=k), K=5)
        print("nv rollout", pol, f"{ms:.3f} ms  {N*T/ms/1e6:.1f} Gsteps/s")
    a = torch.rand((N, 1), device="cuda") * 100
    env.reset(seed=1)
    ms = timeit(lambda k: env.step(a), K=20)
    print("nv step", f"{ms:.3f} ms  {N/ms/1e6:.2f} Gsteps/s  {N*222/ms/1e6:.0f} GB/s alg")
    env.close(); del env, a
if which in ("all", "net"):
    N = 1 << 22
    env = pkg.NetInvMgmtBacklogEnv(num_envs=N, device="cuda:0")
    T = env.num_periods
    ms = timeit(lambda k: env.rollout("constant", seed=6000, episode=k), K=5)
    print("net rollout const", f"{ms:.3f} ms  {N*T/ms/1e6:.2f} Gsteps/s")
    a = torch.rand((N, 11), device="cuda") * 100
    env.reset(seed=1)
    ms = timeit(lambda k: env.step(a), K=20)
    print("net step", f"{ms:.3f} ms  {N/ms/1e6:.3f} Gsteps/s  {N*1598/ms/1e6:.0f} GB/s alg")
    env.close(); del env, a
    G = pkg.synthetic_graph(64)
    N = 1 << 17
    env = pkg.NetInvMgmtMasterEnv(graph=G, backlog=False, num_envs=N, device="cuda:0")
    E = len(env.reorder_links)
    a = torch.rand((N, E), device="cuda") * 100
    env.reset(seed=1)
    ms = timeit(lambda k: env.step(a), K=10)
    print("net64 step", f"{ms:.3f} ms  {N/ms/1e6:.4f} Gsteps/s obs_dim", env.obs_dim)
    ms = timeit(lambda k: env.rollout("constant", seed=6000, episode=k, order_fraction=0.02), K=3)
    print("net64 rollout const", f"{ms:.3f} ms  {N*30/ms/1e6:.3f} Gsteps/s")
