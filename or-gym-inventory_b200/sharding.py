"""Multi-GPU plumbing: one process per GPU, instances sharded by contiguous global-id ranges.

The step path has no inter-GPU traffic (instances are independent; Philox streams are keyed by the GLOBAL
instance id, so results do not depend on the number of ranks).  The only exchange is a sum-allreduce of the
8-element float64 episode-statistics vector at the end of a rollout (torch.distributed: NCCL on GPUs, gloo in
the CPU tests).
"""
import math


def shard_range(global_num_envs, rank, world_size):
    """Contiguous shard of instance ids owned by `rank`: returns (env_offset, num_envs)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    lo = global_num_envs * rank // world_size
    hi = global_num_envs * (rank + 1) // world_size
    return lo, hi - lo


def allreduce_summary(summary, group=None):
    """In-place sum-allreduce of a rollout `summary` tensor ([8] float64) over the process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(summary, op=dist.ReduceOp.SUM, group=group)
    return summary


SUMMARY_FIELDS = ("episodes", "sum_return", "sum_return_sq", "sum_sales", "sum_demand", "sum_unfulfilled",
                  "sum_on_hand", "_pad")


def describe_summary(summary, periods):
    """Batch-level metrics of the reference's evaluation report (benchmark_InvManagementBacklogEnv.py:381-441,
    493-504) from an (all-reduced) summary vector."""
    s = [float(x) for x in summary]
    n = max(s[0], 1.0)
    mean = s[1] / n
    var = max(s[2] / n - mean * mean, 0.0)
    return {"episodes": int(s[0]), "TotalReward_mean": mean, "TotalReward_std": math.sqrt(var),
            "AvgServiceLevel": s[3] / s[4] if s[4] > 0 else float("nan"),
            "TotalStockoutQty_mean": s[5] / n, "AvgEndingInv_mean": s[6] / n / max(periods, 1)}
