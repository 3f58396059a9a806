"""Adapters for the RL libraries the reference's benchmark scripts use.

Neither stable-baselines3 nor ray is installed in this image, so the adapters are duck-typed against the protocols
those libraries call (SB3 `VecEnv`: benchmark_NetInvMgmtLostSalesEnv_sb3_rllib.py:195-247 builds
`DummyVecEnv([lambda: Monitor(Env(**cfg))])`; RLlib: `register_env(name, env_creator)` with
`env_creator(env_config)`, :125-131) and tested against protocol stubs.
"""
import numpy as np


class SB3VecEnvAdapter:
    """stable-baselines3 `VecEnv` protocol on top of a batched env created with autoreset_mode='same_step'.

    * observations / rewards / dones come back as numpy arrays (SB3 is numpy-based), observations cast to
      `obs_dtype` (the reference's FloatObsWrapper, benchmark_InvManagementBacklogEnv_sb3_rllib.py:286-296) and actions
      cast to the env's action dtype (IntActionWrapper, :298-);
    * an env that finishes is reset inside the same step; its last observation is reported as
      infos[i]['terminal_observation'] with infos[i]['TimeLimit.truncated'] = True, like DummyVecEnv does.
    """

    def __init__(self, env, obs_dtype=np.float32):
        if env.autoreset_mode != "same_step":
            raise ValueError("SB3VecEnvAdapter needs a batched env built with autoreset_mode='same_step'")
        self.env = env
        self.num_envs = env.num_envs
        self.observation_space = env.single_observation_space
        self.action_space = env.single_action_space
        self.obs_dtype = np.dtype(obs_dtype)
        self._actions = None
        self._seed = None
        self.render_mode = None

    # -- VecEnv API ----------------------------------------------------------------------------------------------
    def seed(self, seed=None):
        self._seed = seed
        return [None if seed is None else seed + i for i in range(self.num_envs)]

    def reset(self):
        obs, _ = self.env.reset(seed=self._seed)
        self._seed = None
        return obs.cpu().numpy().astype(self.obs_dtype, copy=False)

    def _buffers(self, obs, rew):
        """Pinned host staging (allocated once): one batch of asynchronous copies and ONE synchronisation per step."""
        import torch
        if getattr(self, "_pin", None) is None:
            dev = obs.device
            # observations are cast to obs_dtype ON THE DEVICE (FloatObsWrapper semantics): an int64 observation row of
            # the serial env crosses PCIe as float32, half the bytes
            self._obs_torch_dtype = getattr(torch, self.obs_dtype.name)
            self._pin = dict(act=torch.empty((self.num_envs,) + tuple(self.action_space.shape),
                                             dtype=getattr(torch, np.dtype(self.action_space.dtype).name)).pin_memory(),
                             obs=torch.empty(obs.shape, dtype=self._obs_torch_dtype).pin_memory(),
                             rew=torch.empty(rew.shape, dtype=rew.dtype).pin_memory(),
                             done=torch.empty(self.num_envs, dtype=torch.bool).pin_memory(),
                             trunc=torch.empty(self.num_envs, dtype=torch.bool).pin_memory())
            self._act_dev = torch.empty_like(self._pin["act"], device=dev)
        return self._pin

    def step_async(self, actions):
        self._actions = np.asarray(actions).astype(self.action_space.dtype, copy=False)

    def step_wait(self):
        import torch
        env = self.env
        if getattr(self, "_pin", None) is not None:       # pinned staging exists after the first step
            self._pin["act"].numpy()[...] = self._actions.reshape(self._pin["act"].shape)
            self._act_dev.copy_(self._pin["act"], non_blocking=True)
            obs, rew, term, trunc, info = env.step(self._act_dev)
        else:
            obs, rew, term, trunc, info = env.step(self._actions)
        pin = self._buffers(obs, rew)
        pin["obs"].copy_(obs if obs.dtype == self._obs_torch_dtype else obs.to(self._obs_torch_dtype), non_blocking=True)
        pin["rew"].copy_(rew, non_blocking=True)
        pin["done"].copy_(term | trunc, non_blocking=True)
        pin["trunc"].copy_(trunc, non_blocking=True)
        torch.cuda.current_stream(obs.device).synchronize()
        dones = pin["done"].numpy().copy()
        obs_h = pin["obs"].numpy().copy()
        # infos: one shared empty dict for the instances that go on (a list of N references, no Python loop over the
        # batch); only finished instances get a dict of their own (terminal observation, like DummyVecEnv).  Consumers
        # such as VecMonitor write into the infos of finished instances only.
        infos = [{}] * self.num_envs
        idx = np.flatnonzero(dones)
        if idx.size:
            sel = torch.as_tensor(idx, device=obs.device)
            final = info["final_obs"].index_select(0, sel).cpu().numpy().astype(self.obs_dtype, copy=False)
            tr = pin["trunc"].numpy()
            for k, i in enumerate(idx):
                infos[i] = {"terminal_observation": final[k], "TimeLimit.truncated": bool(tr[i])}
        return obs_h, pin["rew"].numpy().astype(np.float32), dones, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        self.env.close()

    def get_attr(self, attr_name, indices=None):
        n = self.num_envs if indices is None else len(np.atleast_1d(indices))
        return [getattr(self.env, attr_name)] * n

    def set_attr(self, attr_name, value, indices=None):
        setattr(self.env, attr_name, value)

    def env_method(self, method_name, *args, indices=None, **kwargs):
        return [getattr(self.env, method_name)(*args, **kwargs)]

    def env_is_wrapped(self, wrapper_class, indices=None):
        n = self.num_envs if indices is None else len(np.atleast_1d(indices))
        return [False] * n

    def get_images(self):
        return [None] * self.num_envs

    def render(self, mode=None):
        return None


def rllib_env_creator(env_cls, **defaults):
    """`register_env(name, rllib_env_creator(NetInvMgmtLostSalesEnv, num_envs=4096))`: RLlib calls the result with
    its `env_config` dict; batch-level keys (`num_envs`, `device`, `env_offset`, `autoreset_mode`) may be given either
    as defaults here or inside env_config, everything else is passed on as the reference's `env_config`."""
    batch_keys = ("num_envs", "device", "env_offset", "autoreset_mode", "info_level", "wide_state")

    def creator(env_config=None):
        cfg = dict(env_config or {})
        kw = dict(defaults)
        for k in batch_keys:
            if k in cfg:
                kw[k] = cfg.pop(k)
        ctor_keys = ("graph", "num_periods", "backlog")  # network ctor arguments that env_config cannot override (:83-85)
        for k in ctor_keys:
            if k in cfg:
                kw[k] = cfg.pop(k)
        return env_cls(env_config=cfg, **kw) if cfg else env_cls(**kw)

    return creator
