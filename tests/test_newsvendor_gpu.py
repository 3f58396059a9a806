"""GPU parity tests of the newsvendor env (csrc/newsvendor.cu) through the C ABI."""
import numpy as np
import pytest

import or_gym_inventory_b200 as pkg
from helpers import golden_files, ids, load_golden, seq_sum

pytestmark = pytest.mark.gpu
NV = golden_files("newsvendor_")
NV_PP = {"order_up_to": dict(safety_factor=1.0), "classic": dict(safety_factor=1.0), "sS": dict(S_factor=1.2)}


def _torch():
    import torch
    return torch


@pytest.mark.parametrize("path", NV, ids=ids(NV))
def test_step_matches_reference(path):
    torch = _torch()
    g, meta = load_golden(path)
    E = len(g["seeds"])
    env = pkg.NewsvendorEnv(num_envs=E, device="cuda:0", autoreset_mode="disabled", **meta["cfg"])
    # the reference draws its parameters from PCG64; replay them through fixed_params (bit-exact float64)
    obs, _ = env.reset(seed=0, options={"fixed_params": g["params"]})
    assert np.array_equal(obs.cpu().numpy(), g["obs"][:, 0])
    assert np.array_equal(env.export_params().cpu().numpy(), g["params"])
    for t in range(env.step_limit):
        obs, r, term, trunc, info = env.step(torch.from_numpy(g["actions"][:, t]).cuda(),
                                             demand=torch.from_numpy(g["demand"][:, t]).cuda())
        assert np.array_equal(obs.cpu().numpy(), g["obs"][:, t + 1]), t
        assert np.array_equal(r.cpu().numpy(), g["reward"][:, t]), t   # bit-exact, including the float32 branches
        assert np.array_equal(trunc.cpu().numpy(), g["truncated"][:, t])
        assert not term.any()
        parts = torch.stack([info["revenue"], info["purchase_cost"], info["holding_cost"],
                             info["lost_sales_penalty"]], dim=1).cpu().numpy()
        assert np.array_equal(parts, g["parts"][:, t])
    assert env.errors() == 0
    env.close()


@pytest.mark.parametrize("path", NV, ids=ids(NV))
def test_rollout_replay_matches_reference(path):
    g, meta = load_golden(path)
    E = len(g["seeds"])
    env = pkg.NewsvendorEnv(num_envs=E, device="cuda:0", **meta["cfg"])
    out = env.rollout("actions", actions=g["actions"][:, :, 0], demand=g["demand"], fixed_params=g["params"],
                      want=("ep_return", "reward_traj", "final_obs", "stats", "summary"))
    assert np.array_equal(out["reward_traj"].cpu().numpy(), g["reward"])
    assert np.array_equal(out["final_obs"].cpu().numpy(), g["obs"][:, -1])
    ret = out["ep_return"].cpu().numpy()
    for e in range(E):
        assert ret[e] == seq_sum(g["reward"][e])
    assert np.array_equal(out["stats"][:, 1].cpu().numpy(), g["demand"].sum(axis=1).astype(np.float64))
    if meta["policy"] in NV_PP:  # the on-device drivers reproduce the reference agents' actions
        out2 = env.rollout(meta["policy"], demand=g["demand"], fixed_params=g["params"],
                           want=("action_traj", "reward_traj"), **NV_PP[meta["policy"]])
        assert np.array_equal(out2["action_traj"].cpu().numpy(), g["actions"][:, :, 0])
        assert np.array_equal(out2["reward_traj"].cpu().numpy(), g["reward"])
    env.close()


def test_large_batch_vs_oracle():
    from oracle import oracle
    torch = _torch()
    N = 10007
    env = pkg.NewsvendorEnv(num_envs=N, device="cuda:0")
    T = env.step_limit
    obs, _ = env.reset(seed=2000)
    par = env.export_params().cpu().numpy()
    assert (par[:, 0] >= 1).all() and (par[:, 0] <= 100).all() and (par[:, 1] <= par[:, 0]).all()
    assert (par[:, 2] <= np.minimum(par[:, 1], 5)).all() and (par[:, 4] < 200).all()
    rng = np.random.default_rng(0)
    acts = (rng.uniform(0, 1, size=(N, T)) * (2 * par[:, 4:5] + 1)).astype(np.float32)
    acts[rng.random((N, T)) < 0.2] = 0.0
    a_d = torch.from_numpy(acts).cuda()
    rew = torch.zeros((N, T), dtype=torch.float64, device="cuda")
    dem = torch.zeros((N, T), dtype=torch.int64, device="cuda")
    for t in range(T):
        obs, r, _, trunc, info = env.step(a_d[:, t])
        rew[:, t] = r
        dem[:, t] = info["demand"]
    last = obs.clone()
    # fused rollout draws the same parameters and the same demand stream
    out = env.rollout("actions", actions=a_d, seed=2000, want=("ep_return", "reward_traj", "final_obs", "stats"))
    assert torch.equal(out["reward_traj"], rew)
    assert torch.equal(out["final_obs"], last)
    assert torch.equal(out["stats"][:, 1], dem.sum(dim=1).double())
    rew, dem = rew.cpu().numpy(), dem.cpu().numpy()
    for e in list(range(0, N, 499)) + [N - 1]:
        o = oracle.newsvendor_episode(env.params, actions=acts[e], demand=dem[e], fixed=par[e])
        assert np.array_equal(o["reward"], rew[e])
        assert np.array_equal(o["obs"][-1], last[e].cpu().numpy())
    env.close()


def test_policies_vs_oracle_large():
    """classic / order-up-to / (s,S) on-device drivers on 5k random instances vs the oracle (Poisson ppf included)."""
    from oracle import oracle
    torch = _torch()
    N = 5000
    env = pkg.NewsvendorEnv(num_envs=N, device="cuda:0")
    env.reset(seed=77)
    par = env.export_params().cpu().numpy()
    T = env.step_limit
    # device demand stream for (seed=77, env, t)
    dem = torch.zeros((N, T), dtype=torch.int64, device="cuda")
    z = torch.zeros(N, dtype=torch.float32, device="cuda")
    for t in range(T):
        _, _, _, _, info = env.step(z)
        dem[:, t] = info["demand"]
    dem_h = dem.cpu().numpy()
    for pol, pp, kw in (("classic", 1.0, dict(safety_factor=1.0)), ("order_up_to", 1.1, dict(safety_factor=1.1)),
                        ("sS", 1.2, dict(S_factor=1.2))):
        out = env.rollout(pol, seed=77, want=("ep_return", "action_traj", "stats", "summary"), **kw)
        acts, ret = out["action_traj"].cpu().numpy(), out["ep_return"].cpu().numpy()
        for e in range(0, N, 97):
            o = oracle.newsvendor_episode(env.params, policy=pol, pparam=pp, demand=dem_h[e], fixed=par[e])
            assert np.array_equal(o["actions"], acts[e]), (pol, e)
            assert ret[e] == seq_sum(o["reward"])
        s = out["summary"].cpu().numpy()
        assert s[0] == N and np.isclose(s[1], ret.sum(), rtol=1e-12)
    env.close()


def test_autoreset_next_step_draws_new_parameters():
    torch = _torch()
    env = pkg.NewsvendorEnv(num_envs=200, device="cuda:0", step_limit=3)
    obs0 = env.reset(seed=5)[0].clone()
    a = torch.full((200, 1), 10.0, device="cuda")
    for _ in range(3):
        obs, r, term, trunc, _ = env.step(a)
    assert trunc.all()
    obs, r, term, trunc, _ = env.step(a)       # autoreset step
    assert (r == 0).all() and not trunc.any() and (obs[:, 5:] == 0).all()
    assert not torch.equal(obs[:, :5], obs0[:, :5])   # a new episode draws new economics
    env.close()


@pytest.mark.parametrize("mode", ["next_step", "same_step"])
def test_fixed_params_survive_autoreset(mode):
    """reset(options={'fixed_params': ...}) pins the instance for every later episode of the step API
    (CustomizableNewsvendorEnv re-applies them on each reset, benchmark_newsvendor_sb3_rllib.py:276-291); an env reset
    without them keeps drawing fresh parameters per episode.  The pinned demand stream still advances per episode."""
    torch = _torch()
    N, T = 70, 6
    fixed = dict(price=50.0, cost=30.0, h=2.5, k=7.25, mu=45.5)
    env = pkg.NewsvendorEnv(num_envs=N, device="cuda:0", autoreset_mode=mode, step_limit=T)
    a = torch.full((N, 1), 40.0, device="cuda")
    env.reset(seed=11, options={"fixed_params": fixed})
    p0 = env.export_params().cpu().numpy()
    assert np.array_equal(p0, np.tile([50.0, 30.0, 2.5, 7.25, 45.5], (N, 1)))
    dem = []
    for _ in range(3 * T + 3):
        obs, r, term, trunc, info = env.step(a)
        dem.append(info["demand"].cpu().numpy().copy())
        assert np.array_equal(obs[:, :5].cpu().numpy(), p0.astype(np.float32))
    assert np.array_equal(env.export_params().cpu().numpy(), p0)
    stride = T + 1 if mode == "next_step" else T
    assert not np.array_equal(dem[0], dem[stride])          # episode 1 draws its own demand path
    # masked re-reset without fixed_params un-pins exactly those envs
    mask = np.zeros(N, bool)
    mask[::2] = True
    env.reset(options={"reset_mask": mask})
    for _ in range(2 * T + 2):
        env.step(a)
    p1 = env.export_params().cpu().numpy()
    assert np.array_equal(p1[1::2], p0[1::2]) and not np.array_equal(p1[::2], p0[::2])
    env.close()
    # without pinning: parameters change from episode to episode (reference reset :105-111)
    env = pkg.NewsvendorEnv(num_envs=N, device="cuda:0", autoreset_mode=mode, step_limit=T)
    env.reset(seed=11)
    q0 = env.export_params().cpu().numpy()
    for _ in range(T + 1):
        env.step(a)
    assert not np.array_equal(env.export_params().cpu().numpy(), q0)
    env.close()
