#!/usr/bin/env python
"""Time the UNMODIFIED Python reference in the build container (it cannot run on the GPU box: /root/reference and
gymnasium are absent there).  Single process and a multiprocessing pool of os.cpu_count() workers (the stand-in for
gymnasium.vector.AsyncVectorEnv, which is not installed).  TEST INFRASTRUCTURE / documentation only.

    python oracle/ref_rates.py  ->  prints a JSON dict of env-steps/s (recorded in DESIGN.md)
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
            if kind == "invmgmt":
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
    out = {"cpu_count": os.cpu_count()}
    for kind, secs in (("invmgmt", 4), ("newsvendor", 4), ("netinv", 6)):
        out[kind + "_1proc"] = run(kind, secs, 1000)
        with mp.Pool(os.cpu_count()) as pool:
            out[kind + f"_{os.cpu_count()}proc"] = sum(pool.map(_w, [(kind, secs, 1000 + 100000 * i) for i in range(os.cpu_count())]))
    print(json.dumps(out, indent=1))
