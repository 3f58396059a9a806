#!/usr/bin/env python
"""Python quickstart (needs a B200 and the built library: `python __graft_entry__.py`).

The reference's evaluation loop for the default serial multi-echelon env -- `evaluate_agent` with `BaseStockAgent`
(benchmark_InvManagementBacklogEnv.py:142-198, :381-441) -- for one million instances at once, three ways:
the Gymnasium-style step loop, the fused rollout, and the pipelined evaluate() generator."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import or_gym_inventory_b200 as og

N = 1 << 20
env = og.InvManagementBacklogEnv(num_envs=N, device="cuda:0")           # same kwargs / env_config keys as the reference

# 1. reset / step with any policy that produces a [N, 3] tensor of orders (here: a constant order of 20 units)
obs, info = env.reset(seed=4000)                                        # instance i is seeded with 4000 + i
actions = torch.full((N, 3), 20, dtype=torch.int64, device="cuda")
total = torch.zeros(N, dtype=torch.float64, device="cuda")
for _ in range(env.num_periods):
    obs, reward, terminated, truncated, info = env.step(actions)        # tensors alias the env's buffers (zero-copy)
    total += reward
print("step loop, constant order 20: mean episode return", total.mean().item())

# 2. the whole episode in one launch with the on-device base-stock policy (kernel specialised for this config on first use)
out = env.rollout("base_stock", seed=4000, safety_factor=1.0, want=("ep_return", "stats", "summary"))
print("fused rollout, base-stock:", og.evaluation_report(out, env.num_periods))
print("specialised kernel in use:", env.rollout_specialised)

# 3. many episodes: per batch the reference's summary row is computed on the device (csrc/report.cu) and only those 16
#    numbers are copied to pinned host memory, while the next rollout already runs
for k, res in enumerate(env.evaluate("base_stock", episodes=5, seed=4000, safety_factor=1.0)):
    print("episode batch", k, og.report_to_dict(res["report"]))
#    per-instance results instead (PCIe-bound for large batches): want=("ep_return", "stats", "summary")
for k, res in enumerate(env.evaluate("base_stock", episodes=2, seed=4000, safety_factor=1.0, want=("ep_return",))):
    print("episode batch", k, "mean return", float(res["ep_return"].mean()))
env.close()
