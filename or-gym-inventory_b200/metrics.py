"""The reference's benchmark summary row, computed from the outputs of a fused rollout.

`evaluate_agent` (benchmark_InvManagementBacklogEnv.py:381-441, benchmark_NetInvMgmtBacklogEnv.py:223-303) records per
episode TotalReward, AvgServiceLevel, TotalStockoutQty and AvgEndingInv; `process_and_report_results` (:493-504 / :320-330)
aggregates them per agent with pandas: mean / median / std (ddof = 1) / min / max of TotalReward and the mean of the other
three.  `evaluation_report` reproduces that row for a whole batch of episodes from the `ep_return` and `stats` tensors of
`env.rollout(...)`, on the device the tensors live on; with a process group (one rank per GPU) every statistic --
including the median, by bisection on the order statistic -- is exact over all ranks.

`summary` (8 sums reduced inside the rollout, `sharding.describe_summary`) is the cheap alternative when only means and
the standard deviation are needed.
"""
import ctypes as C
import math

REPORT_LEN = 16
REPORT_FIELDS = ("SuccessfulEpisodes", "AvgReward", "MedianReward", "StdReward", "MinReward", "MaxReward",
                 "AvgServiceLevel", "AvgStockoutQty", "AvgEndInv")
_STATS_KIND = {"torch.int64": 0, "torch.int32": 1, "torch.float64": 2}


def report_to_dict(rep):
    """float64[REPORT_LEN] (device or host) -> the reference's summary columns as Python numbers."""
    v = [float(x) for x in rep.reshape(-1)[:9].tolist()]
    d = dict(zip(REPORT_FIELDS, v))
    d["SuccessfulEpisodes"] = int(d["SuccessfulEpisodes"])
    return d


def evaluation_report_device(out, periods, *, report=None, scratch=None, stream=None):
    """The summary row for ONE rank's batch of episodes, computed by the library's report kernels
    (csrc/report.cu: radix-select median, two-pass std, fixed-order sums) without leaving the device and without a host
    synchronisation.  `out` is the dict of `env.rollout(...)`; returns (report float64[REPORT_LEN] device tensor,
    scratch) -- pass both back in to reuse the buffers.  `report_to_dict` turns it into the reference's columns."""
    import torch
    from . import _capi
    ret = out["ep_return"]
    st = out.get("stats", out.get("stats32"))
    n = ret.numel()
    dev = ret.device
    lib = _capi.lib()
    need = int(lib.orgym_report_scratch_bytes(n))
    if scratch is None or scratch.numel() * 8 < need or scratch.device != dev:
        scratch = torch.empty((need + 7) // 8, dtype=torch.int64, device=dev)
    if report is None:
        report = torch.zeros(REPORT_LEN, dtype=torch.float64, device=dev)
    kind = _STATS_KIND[str(st.dtype)] if st is not None else 0
    s = torch.cuda.current_stream(dev).cuda_stream if stream is None else stream
    _capi.check(lib.orgym_evaluation_report(dev.index, C.c_void_p(ret.data_ptr()),
                                            C.c_void_p(st.data_ptr() if st is not None else None), kind, n, int(periods),
                                            C.c_void_p(scratch.data_ptr()), C.c_void_p(report.data_ptr()), C.c_void_p(s)))
    return report, scratch


def _dist(group):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        return dist
    return None


def _allreduce(t, op, group):
    d = _dist(group)
    if d is not None:
        d.all_reduce(t, op=getattr(d.ReduceOp, op), group=group)
    return t


def _ordered_key(x):
    """float64 -> int64 with the same ordering (IEEE-754 trick: flip the magnitude bits of negative values)."""
    import torch
    b = x.contiguous().view(torch.int64)
    return torch.where(b < 0, b ^ 0x7FFFFFFFFFFFFFFF, b)


def kth_smallest(x, k, group=None):
    """Exact k-th smallest (0-based) of the union of every rank's 1-D float64 tensor `x`, without gathering the data:
    bisection on the order-preserving integer image of the values, one count + all-reduce per step (<= 64 steps)."""
    import torch
    keys = _ordered_key(x.double())
    lo = _allreduce(keys.min().clone() if keys.numel() else torch.tensor(2 ** 62, device=x.device), "MIN", group)
    hi = _allreduce(keys.max().clone() if keys.numel() else torch.tensor(-2 ** 62, device=x.device), "MAX", group)
    lo, hi = int(lo.item()), int(hi.item())
    while lo < hi:                       # smallest key v with count(keys <= v) >= k + 1
        mid = lo + (hi - lo) // 2
        c = _allreduce((keys <= mid).sum(), "SUM", group)
        if int(c.item()) >= k + 1:
            hi = mid
        else:
            lo = mid + 1
    v = torch.tensor([lo], dtype=torch.int64, device=x.device)
    v = torch.where(v < 0, v ^ 0x7FFFFFFFFFFFFFFF, v)
    return float(v.view(torch.float64).item())


def evaluation_report(out, periods, *, group=None):
    """out: dict from `env.rollout(..., want=("ep_return", "stats", ...))` -- `ep_return` float64[N] and `stats` [N,4]
    (sales, demand, unfulfilled / lost, on-hand sum; `stats32` is accepted too).  Returns the reference's summary
    columns AvgReward, MedianReward, StdReward, MinReward, MaxReward, AvgServiceLevel, AvgStockoutQty, AvgEndInv and
    SuccessfulEpisodes.  Newsvendor rollouts (whose reference report has no operational columns) get the reward
    columns plus the same three derived from their (sales, demand, lost sales, excess inventory) statistics."""
    import torch
    ret = out["ep_return"].double().reshape(-1)
    st = out.get("stats", out.get("stats32"))
    dev = ret.device
    n = _allreduce(torch.tensor([ret.numel()], dtype=torch.float64, device=dev), "SUM", group)
    n_tot = int(n.item())
    if n_tot == 0:
        raise ValueError("no episodes")
    s1 = _allreduce(ret.sum().reshape(1), "SUM", group)
    mean = float(s1.item()) / n_tot
    ss = _allreduce(((ret - mean) ** 2).sum().reshape(1), "SUM", group)          # two-pass: stable for large N
    std = math.sqrt(float(ss.item()) / (n_tot - 1)) if n_tot > 1 else float("nan")   # pandas: ddof = 1
    big = torch.finfo(torch.float64).max
    mn = _allreduce((ret.min() if ret.numel() else torch.tensor(big, device=dev)).reshape(1).clone(), "MIN", group)
    mx = _allreduce((ret.max() if ret.numel() else torch.tensor(-big, device=dev)).reshape(1).clone(), "MAX", group)
    if _dist(group) is None:
        srt = torch.sort(ret).values
        med = 0.5 * (float(srt[(n_tot - 1) // 2]) + float(srt[n_tot // 2]))       # pandas median (even N: midpoint)
    else:
        med = 0.5 * (kth_smallest(ret, (n_tot - 1) // 2, group) + kth_smallest(ret, n_tot // 2, group))
    rep = {"AvgReward": mean, "MedianReward": med, "StdReward": std, "MinReward": float(mn.item()),
           "MaxReward": float(mx.item()), "SuccessfulEpisodes": n_tot}
    if st is not None:
        st = st.double().reshape(-1, 4)
        sales, dem, unf, inv = st[:, 0], st[:, 1], st[:, 2], st[:, 3]
        # :425 -- per-episode ratio, 1.0 for an episode without demand; the report averages the ratios
        sl = torch.where(dem > 1e-6, sales / torch.clamp(dem, min=1e-6), torch.ones_like(dem))
        sums = torch.stack([sl.sum(), unf.sum(), (inv / max(int(periods), 1)).sum()])
        sums = _allreduce(sums, "SUM", group)
        rep.update(AvgServiceLevel=float(sums[0]) / n_tot, AvgStockoutQty=float(sums[1]) / n_tot,
                   AvgEndInv=float(sums[2]) / n_tot)
    return rep
