"""SB3 VecEnv / RLlib creator adapters against protocol stubs (the libraries are not installed in this image)."""
import numpy as np
import pytest

import or_gym_inventory_b200 as pkg

pytestmark = pytest.mark.gpu


def test_sb3_vecenv_protocol_same_step_autoreset():
    cfg = dict(periods=3, I0=[10, 10], p=5, r=[3, 2, 1], k=[1, 1, 1], h=[0.5, 0.2], c=[15, 20], L=[1, 2],
               dist_param={"mu": 8})
    env = pkg.InvManagementBacklogEnv(num_envs=5, device="cuda:0", autoreset_mode="same_step", **cfg)
    venv = pkg.SB3VecEnvAdapter(env)                        # what DummyVecEnv([...]) gives SB3
    assert venv.num_envs == 5 and venv.observation_space.shape == (6,) and venv.action_space.shape == (2,)
    venv.seed(3)
    obs0 = venv.reset()
    assert obs0.dtype == np.float32 and obs0.shape == (5, 6)
    a = np.full((5, 2), 7.9)                                # float actions from a policy network -> int64 (truncation)
    for t in range(3):
        venv.step_async(a)
        obs, rew, dones, infos = venv.step_wait()
        assert obs.shape == (5, 6) and rew.dtype == np.float32 and len(infos) == 5
        assert dones.all() == (t == 2)
    assert np.array_equal(obs, obs0)                        # reset observation after the terminal step
    assert all("terminal_observation" in i and i["TimeLimit.truncated"] for i in infos)
    assert infos[0]["terminal_observation"].shape == (6,) and not np.array_equal(infos[0]["terminal_observation"], obs0[0])
    assert venv.get_attr("num_stages") == [3] * 5 and venv.env_is_wrapped(object) == [False] * 5
    with pytest.raises(ValueError):
        pkg.SB3VecEnvAdapter(pkg.InvManagementBacklogEnv(num_envs=2, device="cuda:0"))
    venv.close()


def test_rllib_env_creator_accepts_env_config():
    creator = pkg.rllib_env_creator(pkg.NetInvMgmtLostSalesEnv, num_envs=64, device="cuda:0")
    env = creator({"num_periods": 12, "backlog": False, "alpha": 0.99})   # RLlib: env_creator(env_config)
    assert env.num_envs == 64 and env.num_periods == 12 and env.backlog is False and env.alpha == 0.99
    obs, _ = env.reset(seed=1)
    assert tuple(obs.shape) == (64, env.obs_dim)
    env2 = pkg.rllib_env_creator(pkg.NewsvendorEnv, device="cuda:0")({"num_envs": 8, "lead_time": 3})
    assert env2.num_envs == 8 and env2.lead_time == 3
    env.close(); env2.close()
