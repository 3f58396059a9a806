"""Host-side plumbing of the batched envs: the action / demand arguments may be torch CUDA tensors (zero-copy), CPU
tensors, numpy arrays of any integer / float dtype, non-contiguous views or nested lists; results are identical."""
import numpy as np
import pytest

import or_gym_inventory_b200 as pkg

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    return torch


def test_serial_env_accepts_any_array_like():
    torch = _torch()
    N = 50
    rng = np.random.default_rng(0)
    acts = rng.integers(0, 150, size=(N, 3))
    dem = rng.poisson(20, size=N)
    variants = [
        torch.from_numpy(acts).cuda(),                       # CUDA int64 (zero-copy)
        torch.from_numpy(acts),                              # CPU tensor
        acts.astype(np.int32),                               # numpy int32
        acts.astype(np.float32),                             # numpy float -> truncation path (integers: same result)
        acts.tolist(),                                       # nested lists
        torch.from_numpy(np.ascontiguousarray(acts.T)).cuda().T,   # non-contiguous CUDA view
    ]
    outs = []
    for a in variants:
        env = pkg.InvManagementBacklogEnv(num_envs=N, device="cuda:0")
        env.reset(seed=3)
        obs, r, term, trunc, info = env.step(a, demand=dem)
        outs.append((obs.cpu().numpy().copy(), r.cpu().numpy().copy()))
        env.close()
    for o, r in outs[1:]:
        assert np.array_equal(o, outs[0][0]) and np.array_equal(r, outs[0][1])


def test_seeding_conventions_and_context_manager():
    torch = _torch()
    with pkg.InvManagementLostSalesEnv(num_envs=16, device="cuda:0") as env:
        env.reset(seed=100)
        a = torch.zeros((16, 3), dtype=torch.int64, device="cuda")
        d1 = env.step(a)[4]["demand_realized"].clone()
        env.reset(seed=list(range(100, 116)))                # explicit seed + i list == int seed
        d2 = env.step(a)[4]["demand_realized"].clone()
        assert torch.equal(d1, d2)
        env.reset()                                          # seed=None: same keys, next episode -> new demand
        d3 = env.step(a)[4]["demand_realized"].clone()
        assert not torch.equal(d1, d3)
        with pytest.raises(NotImplementedError):
            env.reset(seed=[5, 9] + list(range(14)))
        assert env.single_observation_space.shape == (33,) and env.observation_space.shape == (16, 33)
        assert env.action_space.shape == (16, 3) and env.single_action_space.dtype == np.int64
    assert env.closed
    with pytest.raises(RuntimeError):
        pkg.NewsvendorEnv(num_envs=4, device="cuda:0").step(np.zeros((4, 1), np.float32))   # step before reset


def test_newsvendor_and_network_accept_numpy():
    N = 40
    nv = pkg.NewsvendorEnv(num_envs=N, device="cuda:0")
    nv.reset(seed=1)
    o1 = nv.step(np.full((N, 1), 30.0, np.float64))[1].cpu().numpy().copy()     # float64 numpy -> float32 actions
    nv.reset(seed=1)
    o2 = nv.step([[30.0]] * N)[1].cpu().numpy().copy()
    assert np.array_equal(o1, o2)
    net = pkg.NetInvMgmtBacklogEnv(num_envs=N, device="cuda:0")
    net.reset(seed=1)
    r1 = net.step(np.full((N, 11), 40, np.int64), demand=np.full((N, 1), 20))[1].cpu().numpy().copy()
    net.reset(seed=1)
    r2 = net.step(np.full((N, 11), 40.0, np.float32), demand=np.full((N, 1), 20.0))[1].cpu().numpy().copy()
    assert np.array_equal(r1, r2)
    nv.close(); net.close()


@pytest.mark.parametrize("family", ["serial", "newsvendor", "network", "network_generic"])
def test_info_tensor_layouts_agree(family):
    """The C ABI writes the two-dimensional info tensors row-major (info_ld = 0) or column-major (info_ld >= N, what the
    Python host uses for coalesced stores); both must hold the same values."""
    torch = _torch()
    N = 333                                   # ragged: not a multiple of the tile size
    rng = np.random.default_rng(5)

    def make():
        if family == "serial":
            return pkg.InvManagementLostSalesEnv(num_envs=N, device="cuda:0"), rng.integers(0, 120, size=(N, 3)), ("sales", "unfulfilled")
        if family == "newsvendor":
            return pkg.NewsvendorEnv(num_envs=N, device="cuda:0"), (rng.random((N, 1)) * 100).astype(np.float32), ("parts",)
        env = pkg.NetInvMgmtLostSalesEnv(num_envs=N, device="cuda:0", specialise=(family == "network"))
        return env, (rng.random((N, len(env.reorder_links))) * 60).astype(np.float32), ("demand", "sales", "profit_node")

    rng = np.random.default_rng(5)
    env_c, acts, names = make()
    rng = np.random.default_rng(5)
    env_r, _, _ = make()
    field = {"sales": "sales", "unfulfilled": "unfulfilled", "parts": "parts", "demand": "demand", "profit_node": "profit"}
    for nm in names:                          # give env_r plain row-major buffers
        t = torch.full(tuple(env_r._info_t[nm].shape), -7, dtype=env_r._info_t[nm].dtype, device="cuda")
        assert t.is_contiguous() and env_c._info_t[nm].stride(0) == 1      # column-major storage behind [N, dim]
        env_r._info_t[nm] = t
        setattr(env_r._info, field[nm], t.data_ptr())
    env_r._info.info_ld = 0
    env_c.reset(seed=21)
    env_r.reset(seed=21)
    for _ in range(3):
        env_c.step(acts)
        env_r.step(acts)
        for nm in names:
            a, b = env_c._info_t[nm].cpu().numpy(), env_r._info_t[nm].cpu().numpy()
            assert a.shape == b.shape and np.array_equal(a, b), nm
    env_c.close()
    env_r.close()
