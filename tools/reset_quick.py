"""Time the reset kernels (CUDA events, 20 launches after 3 warm-ups)."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import or_gym_inventory_b200 as og

def t(env, n=20):
    for _ in range(3): env.reset(seed=1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): env.reset()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

out = {}
e = og.InvManagementLostSalesEnv(num_envs=1 << 24); out["invmgmt_2^24_ms"] = t(e); e.close(); del e
e = og.NewsvendorEnv(num_envs=1 << 24); out["newsvendor_2^24_ms"] = t(e); e.close(); del e
e = og.NetInvMgmtBacklogEnv(num_envs=1 << 22); out["netinv_2^22_ms"] = t(e); e.close(); del e
print(json.dumps(out))
