#!/usr/bin/env python
"""Generate the committed golden vectors under tests/golden/ -- TEST INFRASTRUCTURE ONLY.

Runs the UNMODIFIED reference environments from /root/reference (Python, so it
can only run in the build container, not on the GPU box) under the
`oracle/shim/gymnasium` stand-in and records, per case, everything a replay
needs: the actions that were applied, the demand the reference sampled from
its own PCG64 stream, and every per-period state / observation / reward the
reference produced.  The C oracle (oracle/orgym_oracle.c) and the CUDA kernels
are then checked against these files; nothing here is imported by the product.

    python oracle/make_golden.py            # rewrites tests/golden/*.npz

The heuristic policies are restated from the reference's benchmark drivers
(those scripts do not import cleanly: SURVEY.md headline facts):
  base-stock          benchmark_InvManagementBacklogEnv.py:142-198
  order-up-to         benchmark_newsvendor.py:97-111
  classic newsvendor  benchmark_newsvendor.py:113-161
  (s,S)               benchmark_newsvendor_sb3_rllib.py:363-371
  constant order      benchmark_NetInvMgmtBacklogEnv.py:119-136
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("ORGYM_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "shim"))
sys.path.insert(0, REF)

import numpy as np  # noqa: E402
import networkx as nx  # noqa: E402
from scipy.stats import poisson  # noqa: E402

import inventory_management as ref_im  # noqa: E402
import newsvendor as ref_nv  # noqa: E402
import network_management as ref_net  # noqa: E402
import network_management_custom as ref_netc  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


# ----------------------------------------------------------------------------
# policies (drivers named in BASELINE.json configs)
# ----------------------------------------------------------------------------
def base_stock(env, obs, sf=1.0):
    n = env.num_stages - 1
    pos = obs[:n].copy()
    t = env.period
    for i in range(n):
        L = int(env.lead_time[i])
        if L == 0:
            continue
        lo = max(0, t - L)
        if t > 0 and lo < t:
            pos[i] += env.action_log[lo:t, i].sum()
    target = (env.lead_time + 1) * env.dist_param.get("mu", 10) * sf
    q = np.maximum(0, target - pos)
    q = np.clip(q, env.action_space.low, env.action_space.high)
    return q.astype(env.action_space.dtype)


def order_up_to(env, obs, sf=1.0):
    mu = obs[4]
    target = mu * (env.lead_time + 1) * sf
    q = max(0, target - obs[5:].sum())
    q = np.clip(q, env.action_space.low[0], env.action_space.high[0])
    return np.array([q], dtype=env.action_space.dtype)


def classic_newsvendor(env, obs, sf=1.0):
    price, cost, h, k, mu = obs[:5]
    if h + k <= 1e-6 or k < 0 or h < 0:
        q = max(0, mu * (env.lead_time + 1) - obs[5:].sum())
    else:
        cr = k / (h + k)
        eff_mu = mu * (env.lead_time + 1) * sf
        lvl = poisson.ppf(cr, mu=max(1e-6, eff_mu))
        q = max(0, lvl - obs[5:].sum())
    q = np.clip(q, env.action_space.low[0], env.action_space.high[0])
    return np.array([q], dtype=env.action_space.dtype)


def s_S(env, obs, S_factor=1.2):
    price, cost, h, k, mu = obs[:5]
    s_lvl = 0
    if h + k > 1e-6:
        s_lvl = poisson.ppf(np.clip(k / (h + k), 0.001, 0.999), mu=max(1e-6, mu * (env.lead_time + 1)))
    s_level = max(0, s_lvl)
    S_level = s_level * S_factor
    pos = obs[5:].sum()
    q = 0
    if pos < s_level:
        q = max(0, S_level - pos)
    q = np.clip(q, env.action_space.low[0], env.action_space.high[0])
    return np.array([q], dtype=env.action_space.dtype)


# ----------------------------------------------------------------------------
# serial multi-echelon env
# ----------------------------------------------------------------------------
def run_invmgmt(cls, cfg, seeds, policy, rng_seed=0):
    """policy: 'base_stock' | 'random' | 'wild' (floats, negatives, over capacity)."""
    eps = []
    arng = np.random.default_rng(rng_seed)
    for seed in seeds:
        env = cls(**json.loads(json.dumps(cfg)))
        obs, _ = env.reset(seed=seed)
        T, n, m = env.num_periods, env.num_stages - 1, env.num_stages
        rec = dict(actions=np.zeros((T, n), np.float64), obs=np.zeros((T + 1, env.pipeline_length), np.int64),
                   reward=np.zeros(T, np.float64), profit=np.zeros(T, np.float64), truncated=np.zeros(T, np.bool_))
        rec["obs"][0] = obs
        for t in range(T):
            if policy == "base_stock":
                a = base_stock(env, obs)
            elif policy == "random":
                a = arng.integers(0, env.supply_capacity + 1).astype(np.int64)
            elif policy == "wild":
                a = arng.uniform(-0.5, 1.6, size=n) * env.supply_capacity  # float, <0 and >c
            else:
                raise ValueError(policy)
            rec["actions"][t] = a
            obs, r, term, trunc, info = env.step(a)
            assert term is False
            rec["obs"][t + 1] = obs
            rec["reward"][t] = r
            rec["profit"][t] = info["period_profit"]
            rec["truncated"][t] = trunc
        rec.update(I=env.I.copy(), B=env.B.copy(), S=env.S.copy(), LS=env.LS.copy(), R=env.R.copy(),
                   D=env.D.copy(), action_log=env.action_log.copy())
        eps.append(rec)
    out = {k: np.stack([e[k] for e in eps]) for k in eps[0]}
    out["seeds"] = np.asarray(seeds, np.int64)
    return out


def invmgmt_cases():
    small = dict(periods=10, I0=[10, 10], p=5, r=[3, 2, 1], k=[1, 1, 1], h=[0.5, 0.2], c=[15, 20], L=[1, 2],
                 dist_param={"mu": 8})  # the reference's own __main__ config (inventory_management.py:461-471)
    zero_lt = dict(periods=12, I0=[30, 40, 20, 60], p=9.5, r=[7.25, 5.5, 3.125, 2.2, 1.1],
                   k=[0.7, 0.3, 0.2, 0.1, 0.05], h=[0.33, 0.21, 0.11, 0.07], c=[40, 35, 50, 45], L=[0, 3, 0, 2],
                   dist_param={"mu": 17}, alpha=0.9)
    one = dict(periods=8, I0=[25], p=4, r=[2.5, 1.5], k=[0.4, 0.2], h=[0.3], c=[30], L=[2], dist_param={"mu": 12},
               alpha=1.0)
    wide = dict(periods=20, I0=[40, 50, 60, 70, 80, 90, 100, 110], p=30,
                r=[25, 22, 19, 16, 13, 10, 7, 4, 2], k=[0.9, 0.8, 0.7, 0.6, 0.5, 0.4, 0.3, 0.2, 0.1],
                h=[0.45, 0.4, 0.35, 0.3, 0.25, 0.2, 0.15, 0.1], c=[60, 65, 70, 75, 80, 85, 90, 95],
                L=[1, 2, 3, 4, 3, 2, 1, 6], dist_param={"mu": 25}, alpha=0.95)  # m=9 -> 8-way pairwise np.sum
    cases = {}
    B, LS = ref_im.InvManagementBacklogEnv, ref_im.InvManagementLostSalesEnv
    kat = [4000, 4001, 4002, 4003]
    cases["invmgmt_default_backlog_basestock"] = (B, {}, kat, "base_stock")
    cases["invmgmt_default_lost_basestock"] = (LS, {}, kat, "base_stock")
    cases["invmgmt_default_backlog_random"] = (B, {}, [42, 43, 44, 45], "random")
    cases["invmgmt_default_lost_random"] = (LS, {}, [5000, 5001, 5002, 5003], "random")
    cases["invmgmt_default_backlog_wild"] = (B, {}, [7, 8, 9], "wild")
    cases["invmgmt_default_lost_wild"] = (LS, {}, [7, 8, 9], "wild")
    cases["invmgmt_small_backlog_random"] = (B, small, [42, 1, 2], "random")
    cases["invmgmt_small_lost_random"] = (LS, small, [123, 1, 2], "random")
    cases["invmgmt_zerolt_backlog_wild"] = (B, zero_lt, [11, 12, 13], "wild")
    cases["invmgmt_zerolt_lost_basestock"] = (LS, zero_lt, [11, 12, 13], "base_stock")
    cases["invmgmt_onestage_backlog_random"] = (B, one, [3, 4], "random")
    cases["invmgmt_wide_backlog_random"] = (B, wide, [21, 22], "random")
    cases["invmgmt_wide_lost_basestock"] = (LS, wide, [21, 22], "base_stock")
    cases["invmgmt_binomial_backlog_random"] = (B, dict(small, dist=2, dist_param={"n": 30, "p": 0.3}), [5, 6], "random")
    cases["invmgmt_randint_lost_random"] = (LS, dict(small, dist=3, dist_param={"low": 2, "high": 14}), [5, 6], "random")
    cases["invmgmt_geometric_backlog_random"] = (B, dict(small, dist=4, dist_param={"p": 0.12}), [5, 6], "random")
    cases["invmgmt_userD_backlog_random"] = (B, dict(small, dist=5, user_D=[3, 0, 9, 12, 7, 7, 1, 0, 20, 5]), [5, 6], "random")
    return cases


# ----------------------------------------------------------------------------
# newsvendor
# ----------------------------------------------------------------------------
def run_newsvendor(cfg, seeds, policy, rng_seed=0, fixed=None):
    eps = []
    arng = np.random.default_rng(rng_seed)
    for seed in seeds:
        env = ref_nv.NewsvendorEnv(**cfg)
        obs, info = env.reset(seed=seed)
        if fixed is not None:  # CustomizableNewsvendorEnv.reset(options={'fixed_params':…}) semantics
            env.price, env.cost, env.h, env.k, env.mu = (fixed[k] for k in ("price", "cost", "h", "k", "mu"))
            env.state = np.zeros(env.obs_dim, dtype=np.float32)
            env.state[:5] = np.array([env.price, env.cost, env.h, env.k, env.mu], dtype=np.float32)
            obs = env._get_obs()
        T = env.step_limit
        rec = dict(params=np.array([env.price, env.cost, env.h, env.k, env.mu], np.float64),
                   actions=np.zeros((T, 1), np.float32), demand=np.zeros(T, np.int64),
                   obs=np.zeros((T + 1, env.obs_dim), np.float32), reward=np.zeros(T, np.float64),
                   parts=np.zeros((T, 4), np.float64), truncated=np.zeros(T, np.bool_))
        rec["obs"][0] = obs
        for t in range(T):
            if policy == "random":
                a = arng.uniform(0, 2.0 * env.mu + 1.0, size=1).astype(np.float32)
            elif policy == "wild":
                a = arng.uniform(-300, 2600, size=1).astype(np.float32)
            elif policy == "zero_some":
                a = np.array([0.0 if arng.random() < 0.5 else arng.uniform(0, 400)], np.float32)
            elif policy == "order_up_to":
                a = order_up_to(env, obs)
            elif policy == "classic":
                a = classic_newsvendor(env, obs)
            elif policy == "sS":
                a = s_S(env, obs)
            else:
                raise ValueError(policy)
            rec["actions"][t] = a
            obs, r, term, trunc, info = env.step(a)
            rec["obs"][t + 1] = obs
            rec["reward"][t] = r
            rec["demand"][t] = info["demand"]
            rec["parts"][t] = [float(info["revenue"]), float(info["purchase_cost"]), float(info["holding_cost"]),
                               float(info["lost_sales_penalty"])]
            rec["truncated"][t] = trunc
        eps.append(rec)
    out = {k: np.stack([e[k] for e in eps]) for k in eps[0]}
    out["seeds"] = np.asarray(seeds, np.int64)
    return out


def newsvendor_cases():
    s = list(range(2000, 2006))
    cases = {}
    for pol in ("random", "wild", "zero_some", "order_up_to", "classic", "sS"):
        cases[f"newsvendor_default_{pol}"] = ({}, s, pol, None)
    cases["newsvendor_lt0_random"] = (dict(lead_time=0, step_limit=25), s[:3], "random", None)
    cases["newsvendor_lt0_wild"] = (dict(lead_time=0, step_limit=25), s[:3], "wild", None)
    cases["newsvendor_lt1_classic"] = (dict(lead_time=1, step_limit=25), s[:3], "classic", None)
    cases["newsvendor_lt9_random"] = (dict(lead_time=9, step_limit=30), s[:3], "random", None)   # 8-way pairwise fp32 sum
    cases["newsvendor_lt12_orderupto"] = (dict(lead_time=12, step_limit=30, max_inventory=900), s[:3], "order_up_to", None)
    cases["newsvendor_lt17_wild"] = (dict(lead_time=17, step_limit=40, max_inventory=6000), s[:3], "wild", None)
    cases["newsvendor_fixed_classic"] = ({}, s[:3], "classic",
                                         dict(price=50.0, cost=30.0, h=2.5, k=7.25, mu=45.5))
    cases["newsvendor_smallmu_sS"] = (dict(mu_max=8.0, step_limit=30), s[:4], "sS", None)
    return cases


# ----------------------------------------------------------------------------
# network env
# ----------------------------------------------------------------------------
def graph_spec(env):
    """JSON-able description of the topology, in networkx insertion order."""
    g = env.graph
    nodes = [[int(j), {k: float(v) for k, v in g.nodes[j].items()}] for j in g.nodes()]
    edges = []
    for u, v in g.edges():
        d = g.edges[u, v]
        a = {k: float(d[k]) for k in ("L", "p", "g", "b") if k in d}
        if "dist_param" in d:
            a["dist_param"] = {k: float(x) for k, x in d["dist_param"].items()}
        edges.append([int(u), int(v), a])
    return {"nodes": nodes, "edges": edges}


def build_graph(spec, env_for_rng):
    """Rebuild a DiGraph from a spec; Poisson sampler bound to the env's RNG like the default graph."""
    g = nx.DiGraph()
    for j, a in spec["nodes"]:
        g.add_node(j, **a)
    for u, v, a in spec["edges"]:
        a = dict(a)
        if "L" in a:
            a["L"] = int(a["L"])
        if "dist_param" in a:
            a["demand_dist_func"] = (lambda **p: env_for_rng[0].np_random.poisson(**p))
        g.add_edge(u, v, **a)
    return g


YIELD_SPEC = {  # factories with yield < 1, fractional I0, a retailer serving two markets, a factory selling to a market
    "nodes": [[0, {}], [9, {}],
              [1, {"I0": 60.5, "h": 0.04}], [2, {"I0": 75.0, "h": 0.03}], [3, {"I0": 140.0, "h": 0.02}],
              [4, {"I0": 300.0, "C": 55.0, "o": 0.02, "v": 0.9, "h": 0.015}],
              [5, {"I0": 250.0, "C": 70.0, "o": 0.01, "v": 0.75, "h": 0.01}],
              [7, {}], [8, {}]],
    "edges": [[2, 9, {"p": 3.0, "b": 0.3, "dist_param": {"lam": 9}}],
              [1, 0, {"p": 2.5, "b": 0.2, "dist_param": {"lam": 14}}],
              [1, 9, {"p": 2.75, "b": 0.25, "dist_param": {"lam": 6}}],
              [2, 0, {"p": 2.6, "b": 0.15, "dist_param": {"lam": 11}}],
              [3, 1, {"L": 2, "p": 1.5, "g": 0.01}], [3, 2, {"L": 0, "p": 1.4, "g": 0.0}],
              [4, 3, {"L": 4, "p": 0.9, "g": 0.008}], [5, 3, {"L": 3, "p": 0.85, "g": 0.006}],
              [5, 2, {"L": 1, "p": 1.0, "g": 0.004}], [5, 1, {"L": 6, "p": 1.1, "g": 0.003}],
              [7, 4, {"L": 1, "p": 0.1, "g": 0.001}], [8, 4, {"L": 0, "p": 0.12, "g": 0.0}],
              [8, 5, {"L": 2, "p": 0.2, "g": 0.002}]],
}


def run_net(make_env, seeds, policy, rng_seed=0, extra_meta=None):
    eps = []
    arng = np.random.default_rng(rng_seed)
    spec = None
    for seed in seeds:
        env = make_env()
        obs, _ = env.reset(seed=seed)
        if spec is None:
            spec = graph_spec(env)
        T, E, M, J = env.num_periods, len(env.reorder_links), len(env.retail_links), len(env.main_nodes)
        high = env.action_space.high
        rec = dict(actions=np.zeros((T, E), np.float32), obs=np.zeros((T + 1, env.obs_dim), np.float32),
                   reward=np.zeros(T, np.float64), profit=np.zeros(T, np.float64), truncated=np.zeros(T, np.bool_))
        rec["obs"][0] = obs
        const = None
        for t in range(T):
            if policy.startswith("const"):
                if const is None:
                    const = (high.copy() * float(policy[5:])).astype(np.float32)
                a = const
            elif policy.startswith("random"):
                a = arng.uniform(0, float(policy[6:] or 0.12), size=E).astype(np.float32) * high
            elif policy == "wild":
                a = (arng.uniform(-0.05, 0.25, size=E) * high).astype(np.float32)
                a[arng.random(E) < 0.3] = np.float32(arng.integers(0, 60)) + np.float32(0.5)  # banker's-rounding ties
            else:
                raise ValueError(policy)
            rec["actions"][t] = a
            obs, r, term, trunc, info = env.step(a)
            rec["obs"][t + 1] = obs
            rec["reward"][t] = r
            rec["profit"][t] = info["profit_period_undiscounted"]
            rec["truncated"][t] = trunc
        rec.update(X=env.X.values.copy(), Y=env.Y.values.copy(), U=env.U.values.copy(), R=env.R.values.copy(),
                   S=env.S.values.copy(), D=env.D.values.copy(), P=env.P.values.copy())
        eps.append(rec)
    out = {k: np.stack([e[k] for e in eps]) for k in eps[0]}
    out["seeds"] = np.asarray(seeds, np.int64)
    e0 = make_env()
    out["meta"] = np.array(json.dumps(dict(
        graph=spec, num_periods=int(e0.num_periods), backlog=bool(e0.backlog), alpha=float(e0.alpha),
        main_nodes=[int(j) for j in e0.main_nodes], reorder_links=[[int(a), int(b)] for a, b in e0.reorder_links],
        retail_links=[[int(a), int(b)] for a, b in e0.retail_links],
        network_links=[[int(a), int(b)] for a, b in e0.network_links], obs_dim=int(e0.obs_dim),
        action_high=float(e0.action_space.high[0]),
        obs_low=[float(x) for x in e0.observation_space.low], obs_high=[float(x) for x in e0.observation_space.high],
        **(extra_meta or {}))))
    return out


def net_cases():
    cases = {}
    M = ref_net.NetInvMgmtMasterEnv
    cases["net_default_backlog_const0.1"] = (lambda: ref_net.NetInvMgmtBacklogEnv(), [6000, 6001], "const0.1")
    cases["net_default_lostclass_const0.05"] = (lambda: ref_net.NetInvMgmtLostSalesEnv(), [6000], "const0.05")  # runs backlog=True (quirk)
    cases["net_default_backlog_random"] = (lambda: M(backlog=True), [1, 2], "random")
    cases["net_default_lost_random"] = (lambda: M(backlog=False), [1, 2], "random")
    cases["net_default_backlog_wild"] = (lambda: M(backlog=True, alpha=0.96, num_periods=20), [3], "wild")
    cases["net_default_lost_wild"] = (lambda: M(backlog=False, alpha=0.96, num_periods=20), [3], "wild")
    cases["net_custom_lostclass_const0.1"] = (lambda: ref_netc.NetInvMgmtLostSalesEnv(num_periods=40), [7000], "const0.1")
    cases["net_custom_backlog_random"] = (lambda: ref_netc.NetInvMgmtMasterEnv(backlog=True), [8, 9], "random")
    cases["net_custom_lost_random"] = (lambda: ref_netc.NetInvMgmtMasterEnv(backlog=False), [8, 9], "random")

    def mk_yield(backlog):
        def f():
            holder = [None]
            g = build_graph(YIELD_SPEC, holder)
            env = M(graph=g, backlog=backlog, num_periods=16, alpha=0.98)
            # the reference copies the graph (network_management.py:77) -> rebind samplers to THIS env's RNG
            holder[0] = env
            return env
        return f
    cases["net_yield_backlog_random"] = (mk_yield(True), [4, 5], "random")
    cases["net_yield_lost_wild"] = (mk_yield(False), [4, 5], "wild")

    # user_D / sample_path (network_management.py:240-267): a recorded trace drives the market link unless
    # sample_path is set, in which case the link samples from its demand_dist_func as usual.  Fractional trace
    # entries exercise the max(0, int(round(.))) of step (:540); the trace repeats its last entry past its end (:255).
    trace = [17, 0, 23.5, 8, 31, 4.5, 12, 12, 0, 40, 19, 22, 6, 2.5, 27, 15, 9, 33, 18, 1, 0, 21, 14, 29, 7, 11, 25, 3, 16, 20]
    ud = [[1, 0, trace]]

    def mk_user(backlog, sp):
        return lambda: M(backlog=backlog, user_D={(1, 0): list(trace)}, sample_path={(1, 0): sp})
    cases["net_userD_backlog_random"] = (mk_user(True, False), [31, 32], "random", dict(user_D=ud, sample_path=[[1, 0, False]]))
    cases["net_userD_lost_const0.1"] = (mk_user(False, False), [33], "const0.1", dict(user_D=ud, sample_path=[[1, 0, False]]))
    cases["net_userD_samplepath_backlog_random"] = (mk_user(True, True), [34, 35], "random",
                                                    dict(user_D=ud, sample_path=[[1, 0, True]]))

    def mk_yield_user():  # four retail links: two replay traces (one of them with sample_path -> sampled), two sample
        holder = [None]
        g = build_graph(YIELD_SPEC, holder)
        env = M(graph=g, backlog=True, num_periods=16, alpha=0.98,
                user_D={(2, 9): list(trace[:16]), (1, 9): list(trace[3:19])}, sample_path={(2, 9): False, (1, 9): True})
        holder[0] = env
        return env
    cases["net_yield_userD_backlog_wild"] = (mk_yield_user, [36, 37], "wild",
                                             dict(user_D=[[2, 9, trace[:16]], [1, 9, trace[3:19]]],
                                                  sample_path=[[2, 9, False], [1, 9, True]]))

    # BASELINE config 5's synthetic 64-node network, built by the product's own generator and handed to the reference
    # as a networkx graph (classification :146-195 and the step at 88 reorder links / 44 nodes)
    sys.path.insert(0, os.path.dirname(HERE))
    from or_gym_inventory_b200.network_management import synthetic_graph

    def spec_of(g):
        nodes = [[int(j), {k: float(v) for k, v in g.nodes[j].items()}] for j in g.nodes()]
        edges = []
        for u, v in g.edges():
            d = g.edges[u, v]
            a = {k: float(d[k]) for k in ("L", "p", "g", "b") if k in d}
            if "dist_param" in d:
                a["dist_param"] = {k: float(x) for k, x in d["dist_param"].items()}
            edges.append([int(u), int(v), a])
        return {"nodes": nodes, "edges": edges}

    def mk_synth(backlog):
        def f():
            holder = [None]
            g = build_graph(spec_of(synthetic_graph(64)), holder)   # Poisson samplers bound to the env's own RNG
            env = M(graph=g, backlog=backlog, num_periods=30)
            holder[0] = env
            return env
        return f
    cases["net_synth64_lost_random"] = (mk_synth(False), [12000, 12001], "random0.08", dict(synthetic_graph_seed=64))
    cases["net_synth64_backlog_const0.05"] = (mk_synth(True), [12002], "const0.05", dict(synthetic_graph_seed=64))
    return cases


def save(name, arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    return os.path.getsize(path)


def main():
    total = 0
    for name, (cls, cfg, seeds, pol) in invmgmt_cases().items():
        out = run_invmgmt(cls, cfg, seeds, pol)
        out["meta"] = np.array(json.dumps(dict(cfg=cfg, backlog=(cls is ref_im.InvManagementBacklogEnv), policy=pol)))
        total += save(name, out)
    for name, (cfg, seeds, pol, fixed) in newsvendor_cases().items():
        out = run_newsvendor(cfg, seeds, pol, fixed=fixed)
        out["meta"] = np.array(json.dumps(dict(cfg=cfg, policy=pol, fixed=fixed)))
        total += save(name, out)
    for name, (mk, seeds, pol, *extra) in net_cases().items():
        total += save(name, run_net(mk, seeds, pol, extra_meta=extra[0] if extra else None))
    # known-answer values quoted in SURVEY.md §8c (sanity: this script reproduces them)
    g = np.load(os.path.join(OUT, "invmgmt_default_backlog_basestock.npz"))
    assert g["D"][0].tolist() == [19, 16, 16, 27, 25, 17, 16, 11, 19, 26, 20, 31, 24, 22, 19, 19, 16, 15, 23, 19, 20,
                                  16, 15, 26, 23, 12, 19, 11, 22, 25]
    # SURVEY quotes 3891.9482044716797 (its own accumulation order); sequential Python sum gives ...806
    assert abs(float(sum(g["reward"][0].tolist())) - 3891.9482044716797) < 1e-11
    assert g["I"][0][30].tolist() == [0, 150, 0] and g["B"][0][30].tolist() == [7, 0, 0, 0]
    g = np.load(os.path.join(OUT, "net_default_backlog_const0.1.npz"))
    assert g["X"][0][30].tolist() == [4947, 1900, 170, 170, 7340, 680]
    assert abs(float(sum(g["reward"][0].tolist())) + 5901.330000000001) < 1e-10
    print(f"wrote goldens to {OUT}: {total/1024:.0f} KiB")


if __name__ == "__main__":
    main()
