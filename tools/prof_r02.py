"""small launch sequences for the round-2 ncu captures (one target per invocation):
    python tools/prof_r02.py inv_lost | inv_backlog | inv_random | inv_wide | inv_step | nv | nv_step | net | net_step | net64 | report"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import or_gym_inventory_b200 as pkg
which = sys.argv[1]
if which == "report":   # the evaluation-report kernels (csrc/report.cu) on one rollout's outputs, twice
    env = pkg.InvManagementLostSalesEnv(num_envs=1 << 24, device="cuda:0")
    out = env.rollout("base_stock", seed=5000, safety_factor=1.0, want=("ep_return", "stats32"))
    rep, scr = pkg.evaluation_report_device(out, 30)
    rep, scr = pkg.evaluation_report_device(out, 30, report=rep, scratch=scr)
    torch.cuda.synchronize()
    print("ok", pkg.report_to_dict(rep)["MedianReward"])
    sys.exit(0)
if which.startswith("inv"):
    N = 1 << 24
    cls = pkg.InvManagementBacklogEnv if which == "inv_backlog" else pkg.InvManagementLostSalesEnv
    env = cls(num_envs=N, device="cuda:0", wide_state=(which == "inv_wide"))
    if which == "inv_step":
        a = torch.randint(0, 100, (N, 3), dtype=torch.int64, device="cuda")
        env.reset(seed=1)
        for _ in range(14): env.step(a)
    else:
        pol = "random" if which == "inv_random" else "base_stock"
        for k in range(3): env.rollout(pol, seed=5000, episode=k)
elif which in ("nv", "nv_step"):
    N = 1 << 22
    env = pkg.NewsvendorEnv(num_envs=N, device="cuda:0")
    if which == "nv_step":
        a = torch.rand((N, 1), device="cuda") * 100
        env.reset(seed=1)
        for _ in range(8): env.step(a)
    else:
        for k in range(2): env.rollout("classic", seed=2000, episode=k)
elif which in ("net", "net_step"):
    N = 1 << 20
    env = pkg.NetInvMgmtBacklogEnv(num_envs=N, device="cuda:0")
    if which == "net_step":
        a = torch.rand((N, 11), device="cuda") * 100
        env.reset(seed=1)
        for _ in range(12): env.step(a)
    else:
        for k in range(2): env.rollout("constant", seed=6000, episode=k)
elif which == "net64":
    N = 1 << 17
    env = pkg.NetInvMgmtMasterEnv(graph=pkg.synthetic_graph(64), backlog=False, num_envs=N, device="cuda:0",
                                  info_level=int(os.environ.get("INFO", 0)))
    a = torch.rand((N, len(env.reorder_links)), device="cuda") * 100
    env.reset(seed=1)
    for _ in range(8): env.step(a)
torch.cuda.synchronize()
print("ok")
