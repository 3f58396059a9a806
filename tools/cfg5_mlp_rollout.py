#!/usr/bin/env python
"""BASELINE.json config 5 as a runnable measurement: NetInvMgmtLostSalesEnv(graph=G64, backlog=False), 2^17 instances
per GPU, driven through the STEP API by a torch MLP (obs -> 64 -> 64 -> actions, tanh), PPO-style collection of
(obs, action, reward) for one 30-period episode.  Prints one JSON line.  One process per GPU under torchrun; there
is no inter-GPU traffic (instances are sharded by env_offset).

    python tools/cfg5_mlp_rollout.py [--envs-per-gpu 131072] [--episodes 5]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import or_gym_inventory_b200 as pkg

ap = argparse.ArgumentParser()
ap.add_argument("--envs-per-gpu", type=int, default=1 << 17)
ap.add_argument("--episodes", type=int, default=5)
args = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N = args.envs_per_gpu
env = pkg.NetInvMgmtLostSalesEnv(graph=pkg.synthetic_graph(64), backlog=False, num_envs=N, device=dev,
                                 env_offset=rank * N, autoreset_mode="disabled", info_level=0)
E, T = len(env.reorder_links), env.num_periods
torch.manual_seed(0)
mlp = torch.nn.Sequential(torch.nn.Linear(env.obs_dim, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh(),
                          torch.nn.Linear(64, E), torch.nn.Sigmoid()).to(dev)
high = torch.from_numpy(env.single_action_space.high).to(dev) * 0.05
obs_buf = torch.empty((T, N, env.obs_dim), dtype=torch.float32, device=dev)
act_buf = torch.empty((T, N, E), dtype=torch.float32, device=dev)
rew_buf = torch.empty((T, N), dtype=torch.float64, device=dev)


def episode(k):
    obs, _ = env.reset(seed=12000 if k == 0 else None)
    with torch.no_grad():
        for t in range(T):
            obs_buf[t].copy_(obs)
            a = mlp(obs * 1e-3) * high
            act_buf[t].copy_(a)
            obs, r, term, trunc, _ = env.step(a)
            rew_buf[t].copy_(r)


for k in range(2):
    episode(k)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(args.episodes):
    episode(2 + k)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    steps = N * T * args.episodes * world
    print(json.dumps({"config": "cfg5: NetInvMgmtLostSalesEnv(G64, backlog=False) + torch MLP policy, step API",
                      "n_gpus": world, "instances_per_gpu": N, "periods": T, "episodes": args.episodes,
                      "env_steps_per_s": steps / (ms.item() * 1e-3), "ms_per_env_step_batch": ms.item() / (T * args.episodes),
                      "specialised_kernels": env.specialised, "mean_reward_last_episode": float(rew_buf.sum(0).mean())}))
if world > 1:
    dist.destroy_process_group()
