"""Shared helpers for the parity tests."""
import glob
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden_files(prefix):
    return sorted(glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def load_golden(path):
    g = np.load(path)
    meta = json.loads(str(g["meta"]))
    return g, meta


def ids(paths):
    return [os.path.basename(p)[:-4] for p in paths]


def seq_sum(x):
    """Python's `total += reward` accumulation order."""
    s = 0.0
    for v in np.asarray(x).tolist():
        s += v
    return s
