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


def bind_host_to_gpu(device_index):
    """Pin the calling process to the CPU cores NVML reports as local to GPU `device_index` (same NUMA node / PCIe root).
    Pinned host buffers allocated afterwards are first touched there, so the per-episode device->host result copies of
    `evaluate()` do not cross the socket interconnect when 8 ranks stream results at once.  Returns the number of cores
    in the new affinity mask, or 0 when NVML is unavailable (nothing is changed then)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:    # CUDA ordinal -> NVML handle through the PCI address (the two enumerations can differ)
            import torch
            pr = torch.cuda.get_device_properties(int(device_index))
            bus = "%08X:%02X:%02X.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:  # noqa: BLE001
            h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        import os
        return len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001 -- an optimisation only
        return 0

