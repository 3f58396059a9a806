"""Host overhead of one step() call at small batch sizes (the regime of SB3/RLlib drivers): microseconds per call."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import or_gym_inventory_b200 as pkg

out = {}
for name, make, act in (
        ("invmgmt", lambda n: pkg.InvManagementBacklogEnv(num_envs=n), lambda e, n: torch.randint(0, 50, (n, 3), device="cuda")),
        ("newsvendor", lambda n: pkg.NewsvendorEnv(num_envs=n), lambda e, n: torch.rand((n, 1), device="cuda") * 50),
        ("netinv", lambda n: pkg.NetInvMgmtBacklogEnv(num_envs=n), lambda e, n: torch.rand((n, len(e.reorder_links)), device="cuda") * 50)):
    for n in (1, 1024, 65536):
        env = make(n)
        a = act(env, n)
        env.reset(seed=1)
        for _ in range(50):
            env.step(a)
        torch.cuda.synchronize()
        K = 2000
        t0 = time.perf_counter()
        for _ in range(K):
            env.step(a)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out[f"{name}_N{n}_us_per_step"] = round(dt / K * 1e6, 1)
        env.close()
print(json.dumps(out))
