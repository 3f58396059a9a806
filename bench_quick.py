"""scratch: quick timing of the serial-env kernels (superseded by bench.py)"""
import sys, time
import torch
import or_gym_inventory_b200 as pkg
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
for cls in (pkg.InvManagementLostSalesEnv, pkg.InvManagementBacklogEnv):
    env = cls(num_envs=N, device="cuda:0")
    T = env.num_periods
    for pol in ("base_stock", "random"):
        for want in (("summary",), ("ep_return", "stats", "summary")):
            for _ in range(3):
                env.rollout(pol, seed=5000, want=want)
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            K = 10
            for k in range(K):
                env.rollout(pol, seed=5000, episode=k, want=want)
            ev[1].record(); torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1]) / K
            print(cls.__name__, pol, want, f"{ms:.3f} ms/rollout  {N*T/ms/1e6:.1f} Gsteps/s")
    # step API
    a = torch.randint(0, 100, (N, 3), dtype=torch.int64, device="cuda")
    env.reset(seed=1)
    for _ in range(3): env.step(a)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    K = 20
    for k in range(K): env.step(a)
    ev[1].record(); torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / K
    print(cls.__name__, "step API", f"{ms:.3f} ms/step  {N/ms/1e6:.2f} Gsteps/s  {N*466/ms/1e6:.0f} GB/s algorithmic")
    env.close()
