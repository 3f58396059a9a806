"""GPU parity tests of the network env (csrc/netinv.cu) through the C ABI."""
import numpy as np
import pytest

import or_gym_inventory_b200 as pkg
from helpers import golden_files, ids, load_golden, net_S_columns, seq_sum

pytestmark = pytest.mark.gpu
NET = golden_files("net_")


def _torch():
    import torch
    return torch


@pytest.fixture(params=["specialised", "stream", "generic"])
def kernel_mode(request, monkeypatch):
    """Run the parity tests through all network kernels: the NVRTC-specialised register-resident one, the
    NVRTC-specialised streaming STEP kernel (meant for large graphs, forced here), and the generic constant-bank
    interpreter."""
    monkeypatch.setenv("ORGYM_NET_JIT", "0" if request.param == "generic" else "2")
    monkeypatch.setenv("ORGYM_NET_JIT_STREAM", "1" if request.param == "stream" else "0")
    return request.param


def _mk(meta, n, **kw):
    return pkg.NetInvMgmtMasterEnv(graph=meta["graph"], num_periods=meta["num_periods"], backlog=meta["backlog"],
                                   alpha=meta["alpha"], num_envs=n, device="cuda:0", **kw)


@pytest.mark.parametrize("path", NET, ids=ids(NET))
def test_step_matches_reference(path, kernel_mode):
    torch = _torch()
    g, meta = load_golden(path)
    E = len(g["seeds"])
    env = _mk(meta, E, autoreset_mode="disabled")
    cols = net_S_columns(meta)
    obs, _ = env.reset(seed=0)
    assert np.array_equal(obs.cpu().numpy(), g["obs"][:, 0])
    for t in range(env.num_periods):
        obs, r, term, trunc, info = env.step(torch.from_numpy(g["actions"][:, t]).cuda(),
                                             demand=torch.from_numpy(g["D"][:, t]).cuda())
        assert np.array_equal(obs.cpu().numpy(), g["obs"][:, t + 1]), t
        assert np.array_equal(r.cpu().numpy(), g["reward"][:, t]), t   # bit-exact float64
        assert np.array_equal(trunc.cpu().numpy(), g["truncated"][:, t])
        assert not term.any()
        X, Y, U, per = env.export_state()
        assert np.array_equal(X.cpu().numpy(), g["X"][:, t + 1])
        assert np.array_equal(Y.cpu().numpy(), g["Y"][:, t + 1])
        assert np.array_equal(U.cpu().numpy(), g["U"][:, t + 1])
        assert np.array_equal(info["sales"].cpu().numpy(), g["S"][:, t][:, cols])
        assert np.array_equal(info["profit_node"].cpu().numpy(), g["P"][:, t])
        assert np.array_equal(info["profit_period_undiscounted"].cpu().numpy(), g["profit"][:, t])
        assert np.array_equal(info["demand"].cpu().numpy(), g["D"][:, t])
    assert env.errors() == 0
    env.close()


@pytest.mark.parametrize("path", NET, ids=ids(NET))
def test_rollout_replay_matches_reference(path, kernel_mode):
    g, meta = load_golden(path)
    E = len(g["seeds"])
    env = _mk(meta, E)
    out = env.rollout("actions", actions=g["actions"], demand=g["D"],
                      want=("ep_return", "stats", "reward_traj", "final_X", "final_Y", "final_U", "summary"))
    assert np.array_equal(out["reward_traj"].cpu().numpy(), g["reward"])
    assert np.array_equal(out["final_X"].cpu().numpy(), g["X"][:, -1])
    assert np.array_equal(out["final_Y"].cpu().numpy(), g["Y"][:, -1])
    assert np.array_equal(out["final_U"].cpu().numpy(), g["U"][:, -1])
    ret = out["ep_return"].cpu().numpy()
    for e in range(E):
        assert ret[e] == seq_sum(g["reward"][e])
    st = out["stats"].cpu().numpy()
    assert np.array_equal(st[:, 1], g["D"].sum(axis=(1, 2)))
    assert np.allclose(st[:, 3], np.maximum(g["X"][:, 1:], 0).sum(axis=(1, 2)), rtol=1e-12)
    out2 = env.rollout("actions", actions=np.ascontiguousarray(g["actions"].transpose(1, 0, 2)),
                       demand=np.ascontiguousarray(g["D"].transpose(1, 0, 2)), time_major=True, want=("ep_return",))
    assert np.array_equal(out2["ep_return"].cpu().numpy(), ret)
    env.close()


def test_class_surface_and_constant_policy(kernel_mode):
    """Class names / defaults of the reference, ConstantOrderAgent rollout vs oracle with device-sampled demand."""
    from oracle import oracle
    torch = _torch()
    N = 3001
    env = pkg.NetInvMgmtBacklogEnv(num_envs=N, device="cuda:0", autoreset_mode="disabled")
    assert env.obs_dim == 68 and env.single_action_space.shape == (11,) and env.single_action_space.high[0] == 1700
    assert pkg.NetInvMgmtLostSalesEnv(num_envs=2, device="cuda:0").backlog is True      # reference quirk
    assert pkg.NetInvMgmtLostSalesEnv(num_envs=2, device="cuda:0", backlog=False).backlog is False
    ec = pkg.network_management_custom.NetInvMgmtLostSalesEnv(num_envs=2, device="cuda:0", num_periods=40)
    assert ec.obs_dim == 12 and ec.single_action_space.high[0] == 3200
    a = (env.single_action_space.high * 0.1).astype(np.float32)
    T = env.num_periods
    env.reset(seed=6000)
    dem = torch.zeros((N, T, 1), dtype=torch.float64, device="cuda")
    rew = torch.zeros((N, T), dtype=torch.float64, device="cuda")
    a_d = torch.from_numpy(np.tile(a, (N, 1))).cuda()
    for t in range(T):
        obs, r, _, trunc, info = env.step(a_d)
        dem[:, t] = info["demand"]
        rew[:, t] = r
    last = obs.clone()
    out = env.rollout("constant", order_fraction=0.1, seed=6000, want=("ep_return", "reward_traj", "final_X", "stats", "summary"))
    assert torch.equal(out["reward_traj"], rew)                 # same Philox stream in step and rollout
    assert torch.equal(out["stats"][:, 1], dem.sum(dim=(1, 2)))
    dem_h, rew_h = dem.cpu().numpy(), rew.cpu().numpy()
    assert 15 < dem_h.mean() < 25
    for e in range(0, N, 250):
        o = oracle.netinv_episode(env.params, actions=a, demand=dem_h[e], constant=True)
        assert np.array_equal(o["reward"], rew_h[e])
        assert np.array_equal(o["obs"][-1], last[e].cpu().numpy())
        assert np.array_equal(o["X"][-1], out["final_X"][e].cpu().numpy())
    s = out["summary"].cpu().numpy()
    assert s[0] == N and np.isclose(s[1], out["ep_return"].sum().item(), rtol=1e-12)
    # sharding invariance
    sub = pkg.NetInvMgmtBacklogEnv(num_envs=100, device="cuda:0", env_offset=500)
    o2 = sub.rollout("constant", order_fraction=0.1, seed=6000, want=("ep_return",))
    assert torch.equal(o2["ep_return"], out["ep_return"][500:600])
    env.close()


@pytest.mark.parametrize("spec", [True, False], ids=["specialised-streaming", "generic"])
def test_synthetic_64_node_network_vs_oracle(spec):
    """Config 5's synthetic 64-node network (lost sales): random actions, device demand, vs the oracle -- through the
    specialised kernels (streaming STEP kernel for large graphs) and the generic one."""
    from oracle import oracle
    torch = _torch()
    G = pkg.synthetic_graph(64)
    N = 300
    env = pkg.NetInvMgmtMasterEnv(graph=G, backlog=False, num_envs=N, device="cuda:0", specialise=spec)
    assert env.specialised == spec
    P = env.params
    E, M, T = len(P.reorder_links), len(P.retail_links), env.num_periods
    assert len(G.nodes) == 64
    rng = np.random.default_rng(5)
    acts = (rng.uniform(0, 0.08, size=(N, T, E)) * env.single_action_space.high).astype(np.float32)
    env.reset(seed=12000)
    dem = torch.zeros((N, T, M), dtype=torch.float64, device="cuda")
    rew = torch.zeros((N, T), dtype=torch.float64, device="cuda")
    a_d = torch.from_numpy(acts).cuda()
    for t in range(T):
        obs, r, _, _, info = env.step(a_d[:, t])
        dem[:, t] = info["demand"]
        rew[:, t] = r
    out = env.rollout("actions", actions=a_d, seed=12000, want=("reward_traj", "final_X"))
    assert torch.equal(out["reward_traj"], rew)
    dem_h, rew_h = dem.cpu().numpy(), rew.cpu().numpy()
    for e in range(0, N, 37):
        o = oracle.netinv_episode(P, actions=acts[e], demand=dem_h[e])
        assert np.array_equal(o["reward"], rew_h[e])
        assert np.array_equal(o["obs"][-1], obs[e].cpu().numpy())
    env.close()


def test_autoreset_next_step(kernel_mode):
    torch = _torch()
    env = pkg.NetInvMgmtBacklogEnv(num_envs=130, device="cuda:0", num_periods=3)
    obs0 = env.reset(seed=1)[0].clone()
    a = torch.full((130, 11), 50.0, device="cuda")
    for _ in range(3):
        obs, r, term, trunc, _ = env.step(a)
    assert trunc.all() and not torch.equal(obs, obs0)
    obs, r, term, trunc, _ = env.step(a)
    assert torch.equal(obs, obs0) and (r == 0).all() and not trunc.any()
    env.close()


def test_mlp_policy_drives_step_api_zero_copy():
    """BASELINE config 5 in miniature: a torch MLP (obs -> 64 -> 64 -> actions, tanh, scaled to the action bound)
    drives the 64-node lost-sales network through the step API; observations and actions never leave the GPU.
    The recorded trajectory is replayed through the oracle."""
    from oracle import oracle
    torch = _torch()
    G = pkg.synthetic_graph(64)
    N = 96
    env = pkg.NetInvMgmtMasterEnv(graph=G, backlog=False, num_envs=N, device="cuda:0", autoreset_mode="disabled")
    E, M, T = len(env.reorder_links), len(env.retail_links), env.num_periods
    torch.manual_seed(0)
    mlp = torch.nn.Sequential(torch.nn.Linear(env.obs_dim, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh(),
                              torch.nn.Linear(64, E), torch.nn.Sigmoid()).cuda()
    high = torch.from_numpy(env.single_action_space.high).cuda()
    obs, _ = env.reset(seed=12000)
    acts = torch.zeros((N, T, E), dtype=torch.float32, device="cuda")
    dem = torch.zeros((N, T, M), dtype=torch.float64, device="cuda")
    rew = torch.zeros((N, T), dtype=torch.float64, device="cuda")
    with torch.no_grad():
        for t in range(T):
            assert obs.data_ptr() == env._obs.data_ptr()           # zero-copy: the env's own buffer
            a = mlp(obs / 1000.0) * high * 0.05
            obs, r, term, trunc, info = env.step(a)
            acts[:, t], dem[:, t], rew[:, t] = a, info["demand"], r
    assert trunc.all()
    a_h, d_h, r_h = acts.cpu().numpy(), dem.cpu().numpy(), rew.cpu().numpy()
    for e in range(0, N, 19):
        o = oracle.netinv_episode(env.params, actions=a_h[e], demand=d_h[e])
        assert np.array_equal(o["reward"], r_h[e])
        assert np.array_equal(o["obs"][-1], obs[e].cpu().numpy())
    env.close()


def test_history_buffers_match_reference_frames():
    """record_history=True: X/Y/U/R/S/D/P history tensors equal the reference's DataFrames."""
    torch = _torch()
    path = [p for p in NET if "yield_backlog_random" in p][0]
    g, meta = load_golden(path)
    env = _mk(meta, len(g["seeds"]), autoreset_mode="disabled", record_history=True)
    cols = net_S_columns(meta)
    env.reset(seed=0)
    for t in range(env.num_periods):
        env.step(torch.from_numpy(g["actions"][:, t]).cuda(), demand=torch.from_numpy(g["D"][:, t]).cuda())
    for name in ("X", "Y", "U", "R", "D", "P"):
        assert np.array_equal(getattr(env, name).cpu().numpy(), g[name]), name
    assert np.array_equal(env.S.cpu().numpy(), g["S"][:, :, cols])
    env.close()
