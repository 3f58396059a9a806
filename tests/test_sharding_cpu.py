"""CPU suite: the N>1 host logic (shard ranges + statistics allreduce) under gloo with world_size 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import or_gym_inventory_b200 as pkg


def test_shard_ranges_tile_the_id_space():
    for n in (1, 7, 1 << 20, (1 << 24) + 3):
        for w in (1, 2, 3, 8):
            spans = [pkg.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (o0, c0), (o1, _) in zip(spans, spans[1:]):
                assert o0 + c0 == o1
    with pytest.raises(ValueError):
        pkg.shard_range(10, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    off, cnt = pkg.shard_range(1001, rank, world)
    ids = np.arange(off, off + cnt, dtype=np.float64)
    ret = np.sin(ids) * 100.0                       # stand-in per-episode returns of this shard
    s = torch.tensor([cnt, ret.sum(), (ret * ret).sum(), 3.0 * cnt, 4.0 * cnt, 1.0 * cnt, 60.0 * cnt, 0.0],
                     dtype=torch.float64)
    pkg.allreduce_summary(s)
    q.put((rank, s.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_summary_allreduce_gloo_world2():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids = np.arange(1001, dtype=np.float64)
    ret = np.sin(ids) * 100.0
    for r in range(2):
        s = res[r]
        assert s[0] == 1001 and np.isclose(s[1], ret.sum(), rtol=1e-12, atol=1e-9) and np.isclose(s[2], (ret * ret).sum())
    d = pkg.describe_summary(res[0], periods=30)
    assert d["episodes"] == 1001 and np.isclose(d["TotalReward_mean"], ret.mean(), atol=1e-9)
    assert np.isclose(d["TotalReward_std"], ret.std(), rtol=1e-9) and d["AvgServiceLevel"] == 0.75
    assert d["AvgEndingInv_mean"] == 2.0


def test_bind_host_to_gpu_is_a_no_op_without_nvml():
    import os
    from or_gym_inventory_b200.sharding import bind_host_to_gpu
    before = os.sched_getaffinity(0)
    n = bind_host_to_gpu(0)
    assert n == 0 or n == len(os.sched_getaffinity(0))
    if n == 0:
        assert os.sched_getaffinity(0) == before


def _episodes(n, seed):
    rng = np.random.default_rng(seed)
    ret = rng.normal(3000.0, 900.0, n)
    ret[rng.random(n) < 0.05] *= -1.0                                  # negative returns exercise the key ordering
    dem = rng.integers(0, 900, n).astype(np.float64)
    dem[rng.random(n) < 0.01] = 0.0                                    # episodes without demand count as 100 % served
    sales = np.floor(dem * rng.random(n))
    return ret, np.stack([sales, dem, dem - sales, rng.integers(0, 9000, n).astype(np.float64)], axis=1)


def _pandas_row(ret, st, periods):
    import pandas as pd
    sl = np.where(st[:, 1] > 1e-6, st[:, 0] / np.maximum(1e-6, st[:, 1]), 1.0)
    df = pd.DataFrame(dict(Agent="a", TotalReward=ret, AvgServiceLevel=sl, TotalStockoutQty=st[:, 2], AvgEndingInv=st[:, 3] / periods))
    g = df.groupby("Agent").agg(AvgReward=("TotalReward", "mean"), MedianReward=("TotalReward", "median"),
                                StdReward=("TotalReward", "std"), MinReward=("TotalReward", "min"),
                                MaxReward=("TotalReward", "max"), AvgServiceLevel=("AvgServiceLevel", "mean"),
                                AvgStockoutQty=("TotalStockoutQty", "mean"), AvgEndInv=("AvgEndingInv", "mean"))
    return g.iloc[0].to_dict()


@pytest.mark.parametrize("n", [1, 2, 1000, 1001])
def test_evaluation_report_matches_the_reference_aggregation(n):
    """metrics.evaluation_report == the pandas aggregation of the reference's report (…BacklogEnv.py:493-504)."""
    from or_gym_inventory_b200.metrics import evaluation_report
    ret, st = _episodes(n, n)
    rep = evaluation_report(dict(ep_return=torch.from_numpy(ret), stats=torch.from_numpy(st)), periods=30)
    want = _pandas_row(ret, st, 30)
    assert rep["SuccessfulEpisodes"] == n
    for k, v in want.items():
        if n == 1 and k == "StdReward":
            assert np.isnan(rep[k]) and np.isnan(v)
        elif k in ("MedianReward", "MinReward", "MaxReward"):
            assert rep[k] == v, k
        else:
            assert np.isclose(rep[k], v, rtol=1e-12, atol=1e-9), k


def _report_worker(rank, world, port, q):
    from or_gym_inventory_b200.metrics import evaluation_report, kth_smallest
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ret, st = _episodes(2001, 7)
    off, cnt = pkg.shard_range(2001, rank, world)
    out = dict(ep_return=torch.from_numpy(ret[off:off + cnt].copy()), stats=torch.from_numpy(st[off:off + cnt].copy()))
    rep = evaluation_report(out, periods=30)
    k17 = kth_smallest(out["ep_return"], 17)
    q.put((rank, rep, k17))
    dist.barrier()
    dist.destroy_process_group()


def test_evaluation_report_is_exact_across_ranks_gloo_world2():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_report_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ret, st = _episodes(2001, 7)
    want = _pandas_row(ret, st, 30)
    for _, rep, k17 in res:
        assert k17 == np.sort(ret)[17]
        assert rep["SuccessfulEpisodes"] == 2001
        for k, v in want.items():
            if k in ("MedianReward", "MinReward", "MaxReward"):
                assert rep[k] == v, k
            else:
                assert np.isclose(rep[k], v, rtol=1e-12, atol=1e-9), k


def test_kth_smallest_handles_ties_signed_zero_and_extremes():
    from or_gym_inventory_b200.metrics import kth_smallest
    rng = np.random.default_rng(3)
    cases = [np.array([5.0]), np.array([2.0, 2.0, 2.0, 2.0]), np.array([-0.0, 0.0, -1e-300, 1e-300, -1e308, 1e308]),
             np.round(rng.normal(0, 3, 500)), rng.normal(0, 1e-200, 300), -np.abs(rng.normal(0, 1e6, 257))]
    for x in cases:
        srt = np.sort(x)
        for k in sorted({0, len(x) // 2, len(x) - 1}):
            assert kth_smallest(torch.from_numpy(x.copy()), k) == srt[k]
