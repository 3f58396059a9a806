"""`gymnasium.utils.seeding.np_random` stand-in: same construction as the real library."""
import numpy as np


def np_random(seed=None):
    seed_seq = np.random.SeedSequence(seed)
    np_seed = seed_seq.entropy
    rng = np.random.Generator(np.random.PCG64(seed_seq))
    return rng, np_seed
