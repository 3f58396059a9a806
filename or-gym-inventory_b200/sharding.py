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


def _gpu_numa_node(device_index):
    """NUMA node of a CUDA device from sysfs (-1 when unknown)."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(int(device_index))
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            return int(f.read().strip())
    except Exception:  # noqa: BLE001
        return -1


def _prefer_memory_node(node):
    """set_mempolicy(MPOL_PREFERRED, {node}) for the calling thread: pages allocated from now on (including the pinned
    result buffers cudaHostAlloc creates) come from `node` even when the CPU cores of that node are not available to
    this process.  Returns True on success."""
    import ctypes
    import platform
    if node < 0 or node >= 64 or platform.machine() != "x86_64":
        return False
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        mask = ctypes.c_ulong(1 << node)
        MPOL_PREFERRED, SYS_set_mempolicy = 1, 238
        return libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(65)) == 0
    except Exception:  # noqa: BLE001
        return False


def bind_host_to_gpu(device_index):
    """Keep a rank's host side next to its GPU: pin the process to the CPU cores NVML reports as local to GPU
    `device_index` and prefer that GPU's NUMA node for memory, so that the pinned result buffers of `evaluate()` are
    allocated there and the per-episode device->host copies of 8 ranks do not funnel through one socket.  Best effort:
    returns the number of cores in the affinity mask afterwards (0 when NVML is unavailable; nothing changes then)."""
    _prefer_memory_node(_gpu_numa_node(device_index))
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
