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


def net_params(meta, pkg):
    """NetInvMgmtParams for a golden network case (graph spec + flags recorded by make_golden.py)."""
    return pkg.NetInvMgmtParams(graph=meta["graph"], num_periods=meta["num_periods"], backlog=meta["backlog"],
                                alpha=meta["alpha"])


def net_S_columns(meta):
    """The reference's S frame is ordered by sorted network_links; ours is reorder links then retail links."""
    nl = [tuple(x) for x in meta["network_links"]]
    re = [tuple(x) for x in meta["reorder_links"]]
    rt = [tuple(x) for x in meta["retail_links"]]
    return [nl.index(e) for e in re] + [nl.index(e) for e in rt]
