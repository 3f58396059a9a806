"""CPU suite: the C oracle (oracle/orgym_oracle.c) against the golden vectors produced by the unmodified
reference (oracle/make_golden.py), bit for bit, plus its numpy-compatible random stream."""
import numpy as np
import pytest

import or_gym_inventory_b200 as pkg
from oracle import oracle
from helpers import golden_files, ids, load_golden, seq_sum

INV = golden_files("invmgmt_")


def test_numpy_stream_is_reproduced():
    for seed in (0, 1, 42, 4000, 2**32 + 5, 2**63 + 11):
        ref = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        assert np.array_equal(oracle.np_random(seed, 64), ref.random(64))
    for seed, lam in ((42, 20.0), (7, 0.5), (8, 9.99), (9, 10.0), (10, 187.3), (11, 0.0)):
        ref = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        assert np.array_equal(oracle.np_poisson(seed, lam, 2000), ref.poisson(lam, 2000))


@pytest.mark.parametrize("path", INV, ids=ids(INV))
def test_invmgmt_oracle_matches_reference(path):
    g, meta = load_golden(path)
    P = pkg.InvManagementParams(backlog=meta["backlog"], **meta["cfg"])
    for e in range(len(g["seeds"])):
        o = oracle.invmgmt_episode(P, actions=g["actions"][e], demand=g["D"][e])
        for k in ("obs", "reward", "profit", "I", "B", "S", "LS", "R", "action_log"):
            assert np.array_equal(o[k], g[k][e]), (k, e)
        if P.dist == 1:  # the oracle also reproduces the reference's own PCG64 Poisson demand from the seed
            o2 = oracle.invmgmt_episode(P, actions=g["actions"][e], seed=int(g["seeds"][e]))
            assert np.array_equal(o2["D"], g["D"][e])
            assert np.array_equal(o2["reward"], g["reward"][e])
        if meta["policy"] == "base_stock":
            o3 = oracle.invmgmt_episode(P, policy="base_stock", demand=g["D"][e])
            assert np.array_equal(o3["action_log"], g["action_log"][e])
            assert np.array_equal(o3["reward"], g["reward"][e])


def test_invmgmt_known_answers():
    """SURVEY.md §8c known-answer values (reference run with seed 4000, base-stock SF=1.0)."""
    P = pkg.InvManagementParams(backlog=True)
    o = oracle.invmgmt_episode(P, policy="base_stock", seed=4000)
    assert o["D"].tolist() == [19, 16, 16, 27, 25, 17, 16, 11, 19, 26, 20, 31, 24, 22, 19, 19, 16, 15, 23, 19, 20, 16,
                               15, 26, 23, 12, 19, 11, 22, 25]
    assert abs(seq_sum(o["reward"]) - 3891.9482044716797) < 1e-11
    assert o["I"][30].tolist() == [0, 150, 0] and o["B"][30].tolist() == [7, 0, 0, 0]
    P = pkg.InvManagementParams(backlog=False)
    o = oracle.invmgmt_episode(P, policy="base_stock", seed=4000)
    assert abs(seq_sum(o["reward"]) - 3704.961267089707) < 1e-11
    assert o["LS"].sum(axis=0).tolist() == [43, 0, 0, 0]
    assert oracle.invmgmt_episode(pkg.InvManagementParams(), actions=np.zeros((30, 3)), seed=42)["obs"][0].tolist() == \
        [100, 150, 200] + [0] * 30


NV = golden_files("newsvendor_")
NV_PP = {"order_up_to": 1.0, "classic": 1.0, "sS": 1.2}


@pytest.mark.parametrize("path", NV, ids=ids(NV))
def test_newsvendor_oracle_matches_reference(path):
    g, meta = load_golden(path)
    P = pkg.NewsvendorParams(**meta["cfg"])
    for e in range(len(g["seeds"])):
        seed = int(g["seeds"][e])
        fixed = None if meta["fixed"] is None else g["params"][e]
        # replay: recorded actions + recorded demand; parameters from the numpy-compatible reset stream
        o = oracle.newsvendor_episode(P, actions=g["actions"][e], demand=g["demand"][e], fixed=fixed, seed=seed)
        assert np.array_equal(o["params"], g["params"][e])
        assert np.array_equal(o["obs"], g["obs"][e])
        assert np.array_equal(o["reward"], g["reward"][e])
        assert np.array_equal(o["parts"], g["parts"][e])
        if fixed is None:  # the whole episode from the seed alone, demand included
            o2 = oracle.newsvendor_episode(P, actions=g["actions"][e], seed=seed)
            assert np.array_equal(o2["demand"], g["demand"][e])
            assert np.array_equal(o2["reward"], g["reward"][e])
        if meta["policy"] in NV_PP:  # the drivers recompute the recorded actions
            o3 = oracle.newsvendor_episode(P, policy=meta["policy"], pparam=NV_PP[meta["policy"]],
                                           demand=g["demand"][e], fixed=fixed, seed=seed)
            assert np.array_equal(o3["actions"], g["actions"][e, :, 0])
            assert np.array_equal(o3["reward"], g["reward"][e])


def test_newsvendor_known_answers():
    """NewsvendorEnv() reset(seed=42), seven steps with action 50.5."""
    P = pkg.NewsvendorParams()
    o = oracle.newsvendor_episode(P, actions=np.full(40, 50.5, np.float32), seed=42)
    assert o["params"].tolist() == [77.39560485559633, 33.967262302690486, 4.292989599556912, 6.973680290593639,
                                    18.835469577529906]
    assert o["demand"][0] == 23 and o["reward"][0] == -1875.741399606182
    assert o["demand"][6] == 25 and o["reward"][6] == 110.07213792142147


def test_poisson_ppf_matches_scipy():
    from scipy.stats import poisson
    rng = np.random.default_rng(1)
    q = np.concatenate([rng.random(3000), [1e-12, 1e-6, 0.001, 0.5, 0.999, 1 - 1e-9]])
    mu = np.concatenate([rng.random(3000) * 1300 + 1e-3, [1e-6, 0.5, 20, 200, 1200, 1200]])
    mine = np.array([oracle.poisson_ppf(a, b) for a, b in zip(q, mu)])
    ref = poisson.ppf(q, mu)
    assert np.array_equal(mine, ref)
    assert oracle.poisson_ppf(0.0, 5.0) == poisson.ppf(0.0, 5.0) == -1.0
    assert np.isinf(oracle.poisson_ppf(1.0, 5.0)) and np.isinf(poisson.ppf(1.0, 5.0))


NET = golden_files("net_")


@pytest.mark.parametrize("path", NET, ids=ids(NET))
def test_netinv_oracle_matches_reference(path):
    from helpers import net_params, net_S_columns
    g, meta = load_golden(path)
    P = net_params(meta, pkg)
    # host-side classification reproduces the reference's (network_management.py:146-195)
    assert [int(j) for j in P.main_nodes] == meta["main_nodes"]
    assert [list(e) for e in P.reorder_links] == meta["reorder_links"]
    assert [list(e) for e in P.retail_links] == meta["retail_links"]
    assert P.obs_dim == meta["obs_dim"]
    obs_space, act_space = P.spaces()
    assert float(act_space.high[0]) == meta["action_high"]
    assert np.array_equal(obs_space.low, np.asarray(meta["obs_low"], np.float32))
    assert np.array_equal(obs_space.high, np.asarray(meta["obs_high"], np.float32))
    cols = net_S_columns(meta)
    for e in range(len(g["seeds"])):
        o = oracle.netinv_episode(P, actions=g["actions"][e], demand=g["D"][e])
        for k in ("obs", "reward", "profit", "X", "Y", "U", "R", "P"):
            assert np.array_equal(o[k], g[k][e]), (k, e)
        assert np.array_equal(o["S"], g["S"][e][:, cols])
        # the reference's own PCG64 Poisson demand, from the seed (one draw per retail link per period)
        o2 = oracle.netinv_episode(P, actions=g["actions"][e], seed=int(g["seeds"][e]))
        assert np.array_equal(o2["D"], g["D"][e])
        assert np.array_equal(o2["reward"], g["reward"][e])


def test_netinv_known_answers():
    """default graph, seed 6000, constant action 170 on all links."""
    P = pkg.NetInvMgmtParams()
    assert P.obs_dim == 68 and len(P.reorder_links) == 11 and P.spaces()[1].high[0] == 1700
    o = oracle.netinv_episode(P, actions=np.full(11, 170, np.float32), seed=6000, constant=True)
    assert o["D"][:, 0].tolist() == [20, 17, 23, 23, 18, 17, 15, 15, 14, 25, 16, 19, 15, 9, 12, 13, 21, 16, 20, 16, 20,
                                     15, 21, 20, 22, 20, 22, 23, 22, 14]
    assert o["X"][30].tolist() == [4947, 1900, 170, 170, 7340, 680]
    assert o["Y"][30].tolist() == [850, 510, 720, 830, 810, 880, 960, 0, 170, 340, 0]
    assert abs(seq_sum(o["reward"]) + 5901.330000000001) < 1e-9
    Pc = pkg.NetInvMgmtParams(num_periods=40, default_graph_kind="custom")
    assert Pc.obs_dim == 12 and Pc.spaces()[1].high[0] == 3200
    oc = oracle.netinv_episode(Pc, actions=np.full(5, 320, np.float32), seed=7000, constant=True)
    assert oc["D"][0].tolist() == [20, 28, 21]
    assert abs(seq_sum(oc["reward"]) + 4220.000000000002) < 1e-9


def test_netinv_lostsales_class_quirk():
    """NetInvMgmtLostSalesEnv() runs backlog dynamics in the reference (ctor arg overrides env_config, :83-85)."""
    P = pkg.NetInvMgmtParams(env_config={"backlog": False})
    assert P.backlog is True
    assert pkg.NetInvMgmtParams(backlog=False).backlog is False


@pytest.mark.parametrize("case", range(40))
def test_value_bounds_hold_under_brute_force(case):
    """orgym_invmgmt_value_bounds (the telescoping bound behind the int32 / int64 rollout decision and the specialiser's
    exactness proof) against simulation: random and adversarial order sequences within [0, c], long lead times, both
    backlog modes -- no on-hand inventory, backlog, fulfilled order, sale or unfulfilled quantity may exceed it."""
    import ctypes as C
    import or_gym_inventory_b200 as pkg
    from or_gym_inventory_b200 import _capi
    from oracle import oracle
    rng = np.random.default_rng(7000 + case)
    n = int(rng.integers(1, 7))
    T = int(rng.integers(3, 61))
    c = rng.integers(1, 400, n)
    cfg = dict(periods=T, I0=rng.integers(0, 300, n).tolist(), p=30.0, r=np.sort(rng.uniform(0.5, 25, n + 1))[::-1].round(2).tolist(),
               k=rng.uniform(0, 1, n + 1).round(2).tolist(), h=rng.uniform(0, 0.5, n).round(2).tolist(), c=c.tolist(),
               L=rng.integers(0, 25, n).tolist(), dist_param={"mu": float(rng.integers(1, 80))})
    P = pkg.InvManagementParams(backlog=bool(case % 2), **cfg)
    keep = []
    ccfg = P.to_c(keep)
    xvar, xsum = C.c_double(0), C.c_double(0)
    pmag = C.c_double(0)
    assert _capi.lib().orgym_invmgmt_value_bounds(C.byref(ccfg), C.byref(xvar), C.byref(xsum), C.byref(pmag)) == 0
    worst, worst_mag = 0.0, 0.0
    up, uc = np.asarray(P.unit_price, np.float64), np.asarray(P.unit_cost, np.float64)
    kc, hc = np.asarray(P.demand_cost, np.float64), np.asarray(P.holding_cost, np.float64)
    patterns = [rng.integers(0, c + 1, size=(T, n)) for _ in range(6)]
    patterns += [np.tile(c, (T, 1)), np.zeros((T, n), np.int64),
                 np.where((np.arange(T)[:, None] // 3) % 2 == 0, c, 0),        # bang-bang orders
                 np.where(np.arange(n)[None, :] % 2 == 0, c, 0) * np.ones((T, 1), np.int64)]
    for a in patterns:
        dem = rng.poisson(cfg["dist_param"]["mu"], size=T).astype(np.int64)
        if rng.random() < 0.3:
            dem[:] = 0                                                            # starve the retailer too
        o = oracle.invmgmt_episode(P, actions=a.astype(np.float64), demand=dem)
        for key in ("I", "B", "R", "S", "LS"):
            worst = max(worst, float(np.abs(o[key]).max()))
        stats = [np.maximum(o["I"][1:], 0).sum(), o["LS"][:, 0].sum(), o["S"][:, 0].sum()]
        assert max(stats) <= xsum.value
        # sum of |terms| of every period's profit (the specialiser's exactness proof bounds it by profit_mag)
        unf = o["B"][1:] if P.backlog else o["LS"]
        inv = np.concatenate([np.maximum(o["I"][1:], 0), np.zeros((T, 1), np.int64)], axis=1)
        mag = (np.abs(o["S"]) * (up + uc) + inv * hc + np.abs(unf) * kc).sum(axis=1)
        worst_mag = max(worst_mag, float(mag.max()))
    assert worst <= xvar.value, (worst, xvar.value, cfg)
    assert worst_mag <= pmag.value, (worst_mag, pmag.value, cfg)
