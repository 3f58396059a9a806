"""Stand-alone access to the device demand samplers (K6) -- used by the statistical tests.

Demand on the step path is drawn from Philox4x32-10 keyed by (seed + global env id) with counter
(period, episode, stream).  Fixed distributions (Poisson / binomial / uniform-int / geometric with one parameter
set per env family or retail link) use a Walker alias table built by the library; the Newsvendor's per-instance
Poisson mean uses PTRS (mu >= 10) / CDF inversion (mu < 10) in float64.
"""
import ctypes as C

from . import _capi

_KINDS = {"poisson": _capi.DIST_POISSON, "binomial": _capi.DIST_BINOMIAL, "randint": _capi.DIST_RANDINT,
          "geometric": _capi.DIST_GEOMETRIC}


def sample_demand(kind, p0, p1=0.0, *, seed=0, num_envs=1, per_env=1, env_offset=0, device="cuda:0"):
    """int64[num_envs, per_env]: sample (env e, period t) of the fixed distribution `kind`."""
    import torch
    dev = torch.device(device)
    out = torch.empty((num_envs, per_env), dtype=torch.int64, device=dev)
    d = _capi.make_dist(_KINDS[kind], p0, p1)
    _capi.check(_capi.lib().orgym_sample_demand(C.byref(d), C.c_uint64(seed), env_offset, num_envs, per_env,
                                                dev.index or 0, C.c_void_p(out.data_ptr()),
                                                C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return out


def sample_poisson_mu(mu, *, seed=0, period=0, env_offset=0):
    """int64 tensor like `mu` (float64 CUDA tensor): element i ~ Poisson(mu[i]) keyed by (seed + env_offset + i, period)."""
    import torch
    mu = mu.contiguous().to(torch.float64)
    out = torch.empty(mu.shape, dtype=torch.int64, device=mu.device)
    _capi.check(_capi.lib().orgym_sample_poisson_mu(C.c_void_p(mu.data_ptr()), C.c_uint64(seed), env_offset,
                                                    mu.numel(), int(period), mu.device.index or 0,
                                                    C.c_void_p(out.data_ptr()),
                                                    C.c_void_p(torch.cuda.current_stream(mu.device).cuda_stream)))
    return out
