#!/usr/bin/env python
"""Time the UNMODIFIED Python reference in the build container (it cannot run on the GPU box: /root/reference and
gymnasium are absent there).  Single process and a multiprocessing pool of os.cpu_count() workers (the stand-in for
gymnasium.vector.AsyncVectorEnv, which is not installed).  TEST INFRASTRUCTURE / documentation only.

    python oracle/ref_rates.py  ->  writes profiles/python_reference_rates.json (env-steps/s; bench.py copies it into
                                    its JSON line as cpu_baseline.python_reference, labelled "build container")
"""
import json
import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "shim"))
sys.path.insert(0, os.environ.get("ORGYM_REFERENCE", "/root/reference"))
sys.path.insert(0, HERE)

import numpy as np  # noqa: E402


def run(kind, seconds, seed0):
    import make_golden as mg  # policies restated from the benchmark drivers
    import inventory_management as im
    import newsvendor as nv
    import network_management as net
    steps, ep = 0, 0
    if kind == "invmgmt":
        env = im.InvManagementLostSalesEnv()
    elif kind == "invmgmt_backlog":
        env = im.InvManagementBacklogEnv()
    elif kind == "netinv64":
        sys.path.insert(0, os.path.dirname(HERE))
        from or_gym_inventory_b200.network_management import synthetic_graph
        holder = [None]
        g = synthetic_graph(64)
        for u, v, d in g.edges(data=True):
            if "dist_param" in d:
                d["demand_dist_func"] = (lambda **p: holder[0].np_random.poisson(**p))
        env = net.NetInvMgmtMasterEnv(graph=g, backlog=False)
        holder[0] = env
        a = (env.action_space.high * 0.05).astype(np.float32)
    elif kind == "newsvendor":
        env = nv.NewsvendorEnv()
    else:
        env = net.NetInvMgmtBacklogEnv()
        a = (env.action_space.high * 0.1).astype(np.float32)
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        obs, _ = env.reset(seed=seed0 + ep)
        trunc = False
        while not trunc:
            if kind in ("invmgmt", "invmgmt_backlog"):
                act = mg.base_stock(env, obs)
            elif kind == "newsvendor":
                act = mg.classic_newsvendor(env, obs)
            else:
                act = a
            obs, r, term, trunc, info = env.step(act)
            steps += 1
        ep += 1
    return steps / (time.perf_counter() - t0)


def _w(args):
    return run(*args)


if __name__ == "__main__":
    import platform
    import pandas
    model = next((ln.split(":", 1)[1].strip() for ln in open("/proc/cpuinfo") if ln.startswith("model name")), "?")
    nc = os.cpu_count()
    out = {"where": "build container (the Python reference and /root/reference cannot travel to the GPU box)",
           "cpu_count": nc, "cpu_model": model, "python": platform.python_version(), "numpy": np.__version__,
           "pandas": pandas.__version__, "unit": "env-steps/s",
           "method": "unmodified reference env under oracle/shim/gymnasium, policy restated from the benchmark drivers; "
                     f"'1proc' = single-process Python loop, 'allcore' = multiprocessing pool of {nc} workers each looping "
                     "its own env (stand-in for gymnasium AsyncVectorEnv, which is not installed)", "rates": {}}
    for kind, secs in (("invmgmt", 4), ("invmgmt_backlog", 4), ("newsvendor", 4), ("netinv", 6), ("netinv64", 8)):
        r1 = run(kind, secs, 1000)
        with mp.Pool(nc) as pool:
            rn = sum(pool.map(_w, [(kind, secs, 1000 + 100000 * i) for i in range(nc)]))
        out["rates"][kind] = {"1proc": r1, "allcore": rn, "cores": nc}
    path = os.path.join(os.path.dirname(HERE), "profiles", "python_reference_rates.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))
