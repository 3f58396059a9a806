"""Minimal stand-in for the `gymnasium` package -- TEST INFRASTRUCTURE ONLY.

`gymnasium` is not installed in this image and there is no network.  The
reference environments (`/root/reference/{newsvendor,inventory_management,
network_management}.py`) only use `gymnasium.Env`, `gymnasium.spaces.Box` and
`gymnasium.utils.seeding.np_random`; this shim provides exactly those, with
the real library's seeding recipe (`Generator(PCG64(SeedSequence(seed)))`,
cf. /root/reference/test.py:5), so the UNMODIFIED reference can be imported by
`oracle/make_golden.py` to produce the committed golden vectors.  It is never
imported by the product package.
"""
import numpy as np

from . import spaces  # noqa: F401
from .utils import seeding  # noqa: F401


class Env:
    metadata = {"render_modes": []}
    observation_space = None
    action_space = None
    _np_random = None

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random, _ = seeding.np_random(None)
        return self._np_random

    @np_random.setter
    def np_random(self, value):
        self._np_random = value

    @property
    def unwrapped(self):
        return self

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self._np_random, _ = seeding.np_random(seed)

    def step(self, action):
        raise NotImplementedError

    def render(self, *a, **k):
        return None

    def close(self):
        pass
