"""ctypes front-end of the CPU oracle (oracle/orgym_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module.  It reuses the product's host-side config flattening (or_gym_inventory_b200.*Params.to_c) so that the
oracle and the CUDA library are fed byte-identical config structs; the arithmetic is independent C code.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
LIB_PATH = os.path.join(HERE, "_build", "liborgym_oracle.so")

import or_gym_inventory_b200 as pkg  # noqa: E402
from or_gym_inventory_b200 import _capi  # noqa: E402

_lib = None


def build(force=False):
    src = os.path.join(HERE, "orgym_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_poisson_ppf.restype = C.c_double
        _lib.orc_poisson_ppf.argtypes = [C.c_double, C.c_double]
        for name in ("orc_invmgmt_bench", "orc_newsvendor_bench", "orc_netinv_bench"):
            getattr(_lib, name).restype = C.c_int64
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def np_poisson(seed, lam, n):
    out = np.zeros(n, np.int64)
    lib().orc_np_poisson(C.c_uint64(seed), C.c_double(lam), C.c_int64(n), _p(out))
    return out


def np_random(seed, n):
    out = np.zeros(n, np.float64)
    lib().orc_np_random(C.c_uint64(seed), C.c_int64(n), _p(out))
    return out


def poisson_ppf(q, mu):
    return lib().orc_poisson_ppf(float(q), float(mu))


# ---------------------------------------------------------------------------------------------------------
def invmgmt_episode(params, *, actions=None, demand=None, seed=0, policy=None, safety_factor=1.0, mu=None):
    """One episode of the serial env.  params: InvManagementParams.  policy None -> `actions` float64[T,n];
    'base_stock' -> on-the-fly driver.  demand int64[T] or None (numpy-compatible Poisson from `seed`)."""
    keep = []
    cfg = params.to_c(keep)
    T, n, m, od = int(params.num_periods), params.num_stages - 1, params.num_stages, params.pipeline_length
    out = dict(obs=np.zeros((T + 1, od), np.int64), reward=np.zeros(T), profit=np.zeros(T),
               I=np.zeros((T + 1, n), np.int64), B=np.zeros((T + 1, m), np.int64), S=np.zeros((T, m), np.int64),
               LS=np.zeros((T, m), np.int64), R=np.zeros((T, n), np.int64), action_log=np.zeros((T, n), np.int64),
               D=np.zeros(T, np.int64))
    pol = 1 if policy == "base_stock" else 0
    pp = np.array([safety_factor, params.dist_param.get("mu", 10) if mu is None else mu, 0, 0], np.float64)
    a = np.ascontiguousarray(np.asarray(actions, np.float64)) if actions is not None else None
    d = np.ascontiguousarray(np.asarray(demand, np.int64)) if demand is not None else None
    rc = lib().orc_invmgmt_episode(C.byref(cfg), pol, _p(pp), _p(a), _p(d), C.c_uint64(seed), _p(out["obs"]),
                                   _p(out["reward"]), _p(out["profit"]), _p(out["I"]), _p(out["B"]), _p(out["S"]),
                                   _p(out["LS"]), _p(out["R"]), _p(out["action_log"]), _p(out["D"]))
    if rc != 0:
        raise RuntimeError(f"orc_invmgmt_episode failed: {rc}")
    return out


def invmgmt_bench(params, policy, episodes, threads, seed0=0, safety_factor=1.0):
    keep = []
    cfg = params.to_c(keep)
    pp = np.array([safety_factor, params.dist_param.get("mu", 10), 0, 0], np.float64)
    rs = C.c_double(0)
    steps = lib().orc_invmgmt_bench(C.byref(cfg), 1 if policy == "base_stock" else 2, _p(pp), C.c_uint64(seed0),
                                    C.c_int64(episodes), C.c_int(threads), C.byref(rs))
    return int(steps), rs.value


# ---------------------------------------------------------------------------------------------------------
NV_POLICIES = {None: 0, "actions": 0, "order_up_to": 1, "classic": 2, "sS": 3}


def newsvendor_episode(params, *, actions=None, demand=None, fixed=None, seed=0, policy=None, pparam=1.0):
    """One newsvendor episode.  fixed None -> the five reset uniforms come from the numpy-compatible stream of
    `seed` (like the reference); demand None -> numpy-compatible Poisson from the same stream."""
    cfg = params.to_c()
    T, od = int(params.step_limit), params.obs_dim
    out = dict(params=np.zeros(5), obs=np.zeros((T + 1, od), np.float32), reward=np.zeros(T),
               parts=np.zeros((T, 4)), demand=np.zeros(T, np.int64), actions=np.zeros(T, np.float32))
    pp = np.array([pparam, 0, 0, 0], np.float64)
    a = np.ascontiguousarray(np.asarray(actions, np.float32).reshape(-1)) if actions is not None else None
    d = np.ascontiguousarray(np.asarray(demand, np.int64)) if demand is not None else None
    f = np.ascontiguousarray(np.asarray(fixed, np.float64)) if fixed is not None else None
    rc = lib().orc_newsvendor_episode(C.byref(cfg), NV_POLICIES[policy], _p(pp), _p(a), _p(d), _p(f), C.c_uint64(seed),
                                      _p(out["params"]), _p(out["obs"]), _p(out["reward"]), _p(out["parts"]),
                                      _p(out["demand"]), _p(out["actions"]))
    if rc != 0:
        raise RuntimeError(f"orc_newsvendor_episode failed: {rc}")
    return out


def newsvendor_bench(params, policy, episodes, threads, seed0=0, pparam=1.0):
    cfg = params.to_c()
    pp = np.array([pparam, 0, 0, 0], np.float64)
    rs = C.c_double(0)
    steps = lib().orc_newsvendor_bench(C.byref(cfg), NV_POLICIES[policy], _p(pp), C.c_uint64(seed0),
                                       C.c_int64(episodes), C.c_int(threads), C.byref(rs))
    return int(steps), rs.value


# ---------------------------------------------------------------------------------------------------------
def netinv_episode(params, *, actions, demand=None, seed=0, constant=False):
    """One episode of the network env.  actions float32 [T,E] (or [E] with constant=True); demand float64 [T,M]
    or None (numpy-compatible Poisson draws from `seed`, one per retail link per period like the reference)."""
    keep = []
    cfg = params.to_c(keep)
    T, J, E, M = int(params.num_periods), len(params.main_nodes), len(params.reorder_links), len(params.retail_links)
    od = params.obs_dim
    out = dict(obs=np.zeros((T + 1, od), np.float32), reward=np.zeros(T), profit=np.zeros(T),
               X=np.zeros((T + 1, J)), Y=np.zeros((T + 1, E)), U=np.zeros((T + 1, M)), R=np.zeros((T, E)),
               S=np.zeros((T, E + M)), D=np.zeros((T, M)), P=np.zeros((T, J)))
    a = np.ascontiguousarray(np.asarray(actions, np.float32))
    d = np.ascontiguousarray(np.asarray(demand, np.float64)) if demand is not None else None
    rc = lib().orc_netinv_episode(C.byref(cfg), 1 if constant else 0, _p(a), _p(d), C.c_uint64(seed), _p(out["obs"]),
                                  _p(out["reward"]), _p(out["profit"]), _p(out["X"]), _p(out["Y"]), _p(out["U"]),
                                  _p(out["R"]), _p(out["S"]), _p(out["D"]), _p(out["P"]))
    if rc != 0:
        raise RuntimeError(f"orc_netinv_episode failed: {rc}")
    return out


def netinv_bench(params, const_action, episodes, threads, seed0=0):
    keep = []
    cfg = params.to_c(keep)
    a = np.ascontiguousarray(np.asarray(const_action, np.float32))
    rs = C.c_double(0)
    steps = lib().orc_netinv_bench(C.byref(cfg), _p(a), C.c_uint64(seed0), C.c_int64(episodes), C.c_int(threads),
                                   C.byref(rs))
    return int(steps), rs.value
