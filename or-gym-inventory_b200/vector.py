"""Common host-side plumbing of the batched envs (a `gymnasium.vector.VectorEnv` look-alike).

One object = N independent instances of one reference env resident on one GPU.  The object owns the
device tensors (state, obs, reward, flags) and overwrites them IN PLACE on every reset/step call: the
tensors returned by `step` alias the env's buffers (zero-copy), so callers that keep history must clone.
PyTorch is used for device memory, streams and torch.distributed only; all arithmetic happens in the
CUDA library behind the C ABI (`_capi`).
"""
import ctypes as C

import numpy as np

from . import _capi
from .spaces import batch_box

try:  # pragma: no cover - gymnasium is not in the build image
    from gymnasium.vector import VectorEnv as _VectorEnvBase  # type: ignore
except Exception:  # noqa: BLE001
    class _VectorEnvBase:  # minimal protocol: attributes + context manager
        metadata = {}
        render_mode = None
        closed = False

        def __enter__(self):
            return self

        def __exit__(self, *a):
            self.close()
            return False

        @property
        def unwrapped(self):
            return self


_AUTORESET = {"disabled": _capi.AUTORESET_DISABLED, "next_step": _capi.AUTORESET_NEXT_STEP,
              "same_step": _capi.AUTORESET_SAME_STEP}


def _torch():
    import torch
    return torch


class BatchedEnv(_VectorEnvBase):
    """Base class: handle life-cycle, device buffers, seeding convention, error polling."""

    _family = None  # "invmgmt" | "newsvendor" | "netinv"
    obs_torch_dtype = None

    def _setup_batch(self, num_envs, device, env_offset, autoreset_mode):
        torch = _torch()
        if autoreset_mode not in _AUTORESET:
            raise ValueError(f"autoreset_mode must be one of {sorted(_AUTORESET)}")
        self.num_envs = int(num_envs)
        if self.num_envs < 1:
            raise ValueError("num_envs must be >= 1")
        self.env_offset = int(env_offset)
        self.autoreset_mode = autoreset_mode
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("or-gym-inventory_b200 runs on CUDA devices only (no CPU fallback); got device=%r" % (device,))
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._h = C.c_void_p()
        self._keep = []
        self._seed = 0
        self._seeded = False
        self.single_observation_space = self.observation_space_single
        self.single_action_space = self.action_space_single
        self.observation_space = batch_box(self.single_observation_space, self.num_envs)
        self.action_space = batch_box(self.single_action_space, self.num_envs)
        self.closed = False

    # -- helpers -------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def _alloc_state(self, nbytes):
        torch = _torch()
        self._state = torch.zeros((int(nbytes) + 7) // 8, dtype=torch.int64, device=self.device)

    @staticmethod
    def _ptr(t):
        return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(None)

    def _to_dev(self, x, dtype, shape=None):
        """Accept a torch tensor (CUDA: zero-copy; CPU: one H2D copy) or anything numpy can wrap."""
        torch = _torch()
        if isinstance(x, torch.Tensor):
            t = x
        else:
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(x)))
        if t.dtype != dtype:
            t = t.to(dtype)
        if t.device != self.device:
            t = t.to(self.device, non_blocking=True)
        t = t.contiguous()
        if shape is not None and tuple(t.shape) != tuple(shape):
            t = t.reshape(shape)
        return t

    def _resolve_seed(self, seed):
        """gymnasium vector seeding: int seed -> env i gets seed + i.  Returns (reseed_flag, base_seed)."""
        if seed is None:
            if not self._seeded:
                self._seed = int(np.random.SeedSequence().generate_state(1, np.uint64)[0] >> 1)
                self._seeded = True
                return 1, self._seed
            return 0, self._seed
        if isinstance(seed, (list, tuple, np.ndarray)):
            seed = np.asarray(seed).reshape(-1)
            if seed.size != self.num_envs or not np.array_equal(seed, seed[0] + np.arange(self.num_envs)):
                raise NotImplementedError("per-env seed lists must follow the seed+i convention")
            seed = int(seed[0])
        self._seed = int(seed) & ((1 << 64) - 1)
        self._seeded = True
        return 1, self._seed

    def errors(self, clear=True):
        """Sticky device-side error bits (synchronises the current stream)."""
        bits = C.c_uint32(0)
        _capi.check(_capi.lib().orgym_errors(self._h, C.byref(bits), int(bool(clear)), self._stream()))
        return int(bits.value)

    def check_errors(self):
        bits = self.errors(clear=True)
        if bits & _capi.ERR_STEP_PAST_END:
            raise IndexError("step() called on an env whose episode has ended (autoreset disabled); "
                             "the reference raises IndexError here (inventory_management.py:267)")
        if bits & _capi.ERR_INT32_RANGE:
            raise OverflowError("a state value left the int32 range; construct the env with wide_state=True")

    def _episode_periods(self):
        P = self.params
        return int(getattr(P, "num_periods", getattr(P, "step_limit", 1)))

    def evaluate(self, policy, episodes=1, *, seed=None, first_episode=0, want=("report",), **policy_kwargs):
        """The reference's `evaluate_agent` + `process_and_report_results` loop (e.g.
        benchmark_InvManagementLostSalesEnv.py:239-302, :493-504) for the whole batch: one fused rollout per episode
        index, yielding a dict of HOST (pinned) tensors per index.

        want=("report",) (default): the reference's summary row -- mean / median / std / min / max of the episode
          returns, mean service level, stock-out quantity and ending inventory -- computed on the device right after
          the rollout (csrc/report.cu); only those REPORT_LEN numbers cross PCIe.  `metrics.report_to_dict(res["report"])`
          names them.  The per-episode tensors stay on the device (`self.last_rollout`).
        want=("ep_return", "stats", "summary", ...): per-instance results copied to the host (PCIe-bound for large
          batches: 8-40 B per episode).

        Rollout k+1 runs on the GPU while the results of rollout k are copied on a second stream; three device and
        three pinned host buffer sets rotate, so a yielded dict stays valid while the NEXT one is being consumed and
        is overwritten only after two more have been yielded."""
        torch = _torch()
        main = torch.cuda.current_stream(self.device)
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(self.device)
            self._host_sets = {}
            self._report_bufs = {}
        NSET = 3
        copy_done = [None] * NSET
        pending = None
        do_report = "report" in want
        rwant = tuple(w for w in want if w != "report")
        if do_report:
            from .metrics import evaluation_report_device
            stats_name = "stats32" if self._family == "invmgmt" else "stats"
            rwant = tuple(dict.fromkeys(("ep_return", stats_name) + rwant))
        for k in range(int(episodes)):
            b = k % NSET
            if copy_done[b] is not None:
                main.wait_event(copy_done[b])          # device buffers of set b are free again
            out = self.rollout(policy, seed=seed if k == 0 else None, episode=first_episode + k, want=rwant,
                               buffer_set=b, **policy_kwargs)
            self.last_rollout = out
            ready = torch.cuda.Event()
            ready.record(main)
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(ready)
                if do_report:
                    # the report kernels (HBM-bound) run on the second stream too: they overlap the next rollout
                    # (issue-bound) instead of delaying it
                    rb = self._report_bufs.setdefault(b, [None, None])
                    rb[0], rb[1] = evaluation_report_device(out, self._episode_periods(), report=rb[0], scratch=rb[1])
                    out = {name: t for name, t in out.items() if name in want}
                    out["report"] = rb[0]
                host = self._host_sets.setdefault((b, tuple(sorted(out))), {})
                for name, t in out.items():
                    if name not in host:
                        host[name] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                    host[name].copy_(t, non_blocking=True)
                done = torch.cuda.Event()
                done.record(self._copy_stream)
            copy_done[b] = done
            if pending is not None:
                pending[1].synchronize()
                yield pending[0]
            pending = (host, done)
        if pending is not None:
            pending[1].synchronize()
            yield pending[0]

    def close(self):
        if not getattr(self, "closed", True) and getattr(self, "_h", None):
            destroy = getattr(_capi.lib(), f"orgym_{self._family}_destroy")
            destroy(self._h)
            self._h = C.c_void_p()
        self.closed = True

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
