"""CPU suite, build container only: the C oracle against the UNMODIFIED reference executed live on random configurations.

The committed golden vectors (tests/golden, oracle/make_golden.py) pin the oracle on 42 fixed cases; here the reference
itself (/root/reference, imported under the gymnasium shim exactly like make_golden.py does) is run on freshly drawn
configurations and action sequences and the oracle must reproduce every array bit for bit, including the reference's own
PCG64 / Poisson demand from the seed.  The GPU box has no /root/reference: the whole module is skipped there (nothing in
the gpu-marked tests, smoke() or bench.py ever reads the reference)."""
import importlib.util
import os

import numpy as np
import pytest

import or_gym_inventory_b200 as pkg
from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ORGYM_REFERENCE", "/root/reference")
if not os.path.exists(os.path.join(REF, "inventory_management.py")):
    pytest.skip("the reference is not available on this box", allow_module_level=True)


@pytest.fixture(scope="module")
def mg():
    spec = importlib.util.spec_from_file_location("orgym_make_golden", os.path.join(ROOT, "oracle", "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)          # imports the reference modules under the shim; writes nothing
    return m


@pytest.mark.parametrize("case", range(24))
def test_serial_env_oracle_equals_live_reference(mg, case):
    rng = np.random.default_rng(31000 + case)
    n = int(rng.integers(1, 10))        # up to 9 stages: m >= 8 exercises numpy's pairwise reward sum
    dist = int(rng.choice([1, 1, 2, 3, 4]))
    dp = {1: {"mu": float(rng.integers(1, 60))}, 2: {"n": int(rng.integers(5, 60)), "p": float(rng.uniform(0.1, 0.9))},
          3: {"low": int(rng.integers(0, 10)), "high": int(rng.integers(10, 50))}, 4: {"p": float(rng.uniform(0.05, 0.5))}}[dist]
    cfg = dict(periods=int(rng.integers(2, 45)), I0=rng.integers(0, 250, n).tolist(), p=float(rng.uniform(5, 50)),
               r=np.sort(rng.uniform(0.2, 30, n + 1))[::-1].tolist(), k=rng.uniform(0, 2, n + 1).tolist(),
               h=rng.uniform(0, 1, n).tolist(), c=rng.integers(1, 300, n).tolist(), L=rng.integers(0, 14, n).tolist(),
               dist=dist, dist_param=dp, alpha=float(rng.uniform(0.7, 1.0)))
    backlog = bool(case % 2)
    cls = mg.ref_im.InvManagementBacklogEnv if backlog else mg.ref_im.InvManagementLostSalesEnv
    policy = ["random", "wild", "base_stock"][case % 3]
    g = mg.run_invmgmt(cls, cfg, [int(rng.integers(0, 2 ** 31)), int(rng.integers(0, 2 ** 31))], policy, rng_seed=case)
    P = pkg.InvManagementParams(backlog=backlog, **cfg)
    for e in range(2):
        o = oracle.invmgmt_episode(P, actions=g["actions"][e], demand=g["D"][e])
        for k in ("obs", "reward", "profit", "I", "B", "S", "LS", "R", "action_log"):
            assert np.array_equal(o[k], g[k][e]), (case, k, e)
        if dist == 1:                    # the reference's own demand stream from the seed
            o2 = oracle.invmgmt_episode(P, actions=g["actions"][e], seed=int(g["seeds"][e]))
            assert np.array_equal(o2["D"], g["D"][e]) and np.array_equal(o2["reward"], g["reward"][e])
        if policy == "base_stock":
            o3 = oracle.invmgmt_episode(P, policy="base_stock", demand=g["D"][e], mu=dp.get("mu", 10))
            assert np.array_equal(o3["action_log"], g["action_log"][e])


@pytest.mark.parametrize("case", range(24))
def test_newsvendor_oracle_equals_live_reference(mg, case):
    rng = np.random.default_rng(32000 + case)
    cfg = dict(lead_time=int(rng.integers(0, 20)), step_limit=int(rng.integers(3, 45)),
               max_inventory=int(rng.integers(100, 6000)), max_order_quantity=int(rng.integers(20, 3000)),
               p_max=float(rng.uniform(5, 300)), h_max=float(rng.uniform(0.2, 12)), k_max=float(rng.uniform(0.5, 25)),
               mu_max=float(rng.uniform(2, 400)))
    policy = ["random", "wild", "zero_some", "order_up_to", "classic", "sS"][case % 6]
    seeds = [int(rng.integers(0, 2 ** 31)), int(rng.integers(0, 2 ** 31))]
    g = mg.run_newsvendor(cfg, seeds, policy, rng_seed=case)
    P = pkg.NewsvendorParams(**cfg)
    for e in range(2):
        o = oracle.newsvendor_episode(P, actions=g["actions"][e], demand=g["demand"][e], seed=seeds[e])
        assert np.array_equal(o["params"], g["params"][e])
        for k in ("obs", "reward", "parts"):
            assert np.array_equal(o[k], g[k][e]), (case, k, e)
        o2 = oracle.newsvendor_episode(P, actions=g["actions"][e], seed=seeds[e])     # demand from the seed as well
        assert np.array_equal(o2["demand"], g["demand"][e]) and np.array_equal(o2["reward"], g["reward"][e])
        if policy in ("order_up_to", "classic", "sS"):
            o3 = oracle.newsvendor_episode(P, policy=policy, pparam={"order_up_to": 1.0, "classic": 1.0, "sS": 1.2}[policy],
                                           demand=g["demand"][e], seed=seeds[e])
            assert np.array_equal(o3["actions"], g["actions"][e, :, 0])


def _random_spec(rng):
    """Random layered supply network in make_golden's spec format: markets / retailers / distributors / factories (some with
    yield < 1) / raw material nodes, shuffled node ids, lead times including 0, retailers serving one or two markets."""
    n_mk, n_rt, n_ds, n_fc, n_raw = 2, int(rng.integers(1, 4)), int(rng.integers(1, 3)), int(rng.integers(1, 4)), 2
    ids = list(rng.permutation(n_mk + n_rt + n_ds + n_fc + n_raw) + 3)
    take = lambda k: [int(ids.pop()) for _ in range(k)]  # noqa: E731
    mk, rt, ds, fc, raw = take(n_mk), take(n_rt), take(n_ds), take(n_fc), take(n_raw)
    nodes = [[j, {}] for j in mk]
    nodes += [[j, {"I0": float(rng.choice([rng.integers(20, 300), rng.uniform(20, 300)])), "h": float(rng.uniform(0, 0.1))}]
              for j in rt + ds]
    nodes += [[j, {"I0": float(rng.integers(50, 400)), "h": float(rng.uniform(0, 0.05)), "C": float(rng.integers(20, 120)),
                   "o": float(rng.uniform(0, 0.05)), "v": float(rng.choice([1.0, rng.uniform(0.5, 1.0)]))}] for j in fc]
    nodes += [[j, {}] for j in raw]
    edges = []
    for r in rt:
        for m in rng.choice(mk, size=int(rng.integers(1, 3)), replace=False):
            edges.append([r, int(m), {"p": float(rng.uniform(2, 9)), "b": float(rng.uniform(0, 0.5)),
                                      "dist_param": {"lam": float(rng.integers(1, 30))}}])
    for lower, upper in ((rt, ds), (ds, fc), (fc, raw)):
        for j in lower:
            for s in rng.choice(upper, size=int(rng.integers(1, len(upper) + 1)), replace=False):
                edges.append([int(s), j, {"L": int(rng.integers(0, 9)), "p": float(rng.uniform(0.1, 2)), "g": float(rng.uniform(0, 0.02))}])
    order = rng.permutation(len(edges))
    return {"nodes": nodes, "edges": [edges[i] for i in order]}


@pytest.mark.parametrize("case", range(16))
def test_network_env_oracle_equals_live_reference(mg, case):
    import json
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import net_params, net_S_columns
    rng = np.random.default_rng(33000 + case)
    spec = _random_spec(rng)
    T, backlog, alpha = int(rng.integers(3, 26)), bool(case % 2), float(rng.uniform(0.9, 1.0))

    def make_env():
        holder = [None]
        g = mg.build_graph(spec, holder)
        env = mg.ref_net.NetInvMgmtMasterEnv(graph=g, backlog=backlog, num_periods=T, alpha=alpha)
        holder[0] = env      # the reference copies the graph: bind the Poisson samplers to THIS env's generator
        return env

    seeds = [int(rng.integers(0, 2 ** 31)), int(rng.integers(0, 2 ** 31))]
    g = mg.run_net(make_env, seeds, ["random", "wild", "const0.08"][case % 3], rng_seed=case)
    meta = json.loads(str(g["meta"]))
    P = net_params(meta, pkg)
    assert [int(j) for j in P.main_nodes] == meta["main_nodes"]
    assert [list(e) for e in P.reorder_links] == meta["reorder_links"]
    assert [list(e) for e in P.retail_links] == meta["retail_links"] and P.obs_dim == meta["obs_dim"]
    cols = net_S_columns(meta)
    for e in range(2):
        o = oracle.netinv_episode(P, actions=g["actions"][e], demand=g["D"][e])
        for k in ("obs", "reward", "profit", "X", "Y", "U", "R", "P"):
            assert np.array_equal(o[k], g[k][e]), (case, k, e)
        assert np.array_equal(o["S"], g["S"][e][:, cols])
        o2 = oracle.netinv_episode(P, actions=g["actions"][e], seed=seeds[e])       # demand from the seed too
        assert np.array_equal(o2["D"], g["D"][e]) and np.array_equal(o2["reward"], g["reward"][e])


def test_committed_golden_vectors_regenerate_identically(mg, tmp_path):
    """tests/golden/*.npz are exactly what oracle/make_golden.py produces from the reference in this container."""
    import glob
    old_out = mg.OUT
    mg.OUT = str(tmp_path)
    try:
        mg.main()
    finally:
        mg.OUT = old_out
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))
    assert len(files) == len(glob.glob(os.path.join(str(tmp_path), "*.npz"))) >= 42
    for f in files:
        a, b = np.load(f, allow_pickle=False), np.load(os.path.join(str(tmp_path), os.path.basename(f)), allow_pickle=False)
        assert set(a.files) == set(b.files), f
        for k in a.files:
            assert np.array_equal(a[k], b[k]), (f, k)
