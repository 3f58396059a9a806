"""CPU suite: the C oracle (oracle/orgym_oracle.c) against the golden vectors produced by the unmodified
reference (oracle/make_golden.py), bit for bit, plus its numpy-compatible random stream."""
import numpy as np
import pytest

import or_gym_inventory_b200 as pkg
from oracle import oracle
from helpers import golden_files, ids, load_golden, seq_sum

INV = golden_files("invmgmt_")


def test_numpy_stream_is_reproduced():
    for seed in (0, 1, 42, 4000, 2**32 + 5, 2**63 + 11):
        ref = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        assert np.array_equal(oracle.np_random(seed, 64), ref.random(64))
    for seed, lam in ((42, 20.0), (7, 0.5), (8, 9.99), (9, 10.0), (10, 187.3), (11, 0.0)):
        ref = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        assert np.array_equal(oracle.np_poisson(seed, lam, 2000), ref.poisson(lam, 2000))


@pytest.mark.parametrize("path", INV, ids=ids(INV))
def test_invmgmt_oracle_matches_reference(path):
    g, meta = load_golden(path)
    P = pkg.InvManagementParams(backlog=meta["backlog"], **meta["cfg"])
    for e in range(len(g["seeds"])):
        o = oracle.invmgmt_episode(P, actions=g["actions"][e], demand=g["D"][e])
        for k in ("obs", "reward", "profit", "I", "B", "S", "LS", "R", "action_log"):
            assert np.array_equal(o[k], g[k][e]), (k, e)
        if P.dist == 1:  # the oracle also reproduces the reference's own PCG64 Poisson demand from the seed
            o2 = oracle.invmgmt_episode(P, actions=g["actions"][e], seed=int(g["seeds"][e]))
            assert np.array_equal(o2["D"], g["D"][e])
            assert np.array_equal(o2["reward"], g["reward"][e])
        if meta["policy"] == "base_stock":
            o3 = oracle.invmgmt_episode(P, policy="base_stock", demand=g["D"][e])
            assert np.array_equal(o3["action_log"], g["action_log"][e])
            assert np.array_equal(o3["reward"], g["reward"][e])


def test_invmgmt_known_answers():
    """SURVEY.md §8c known-answer values (reference run with seed 4000, base-stock SF=1.0)."""
    P = pkg.InvManagementParams(backlog=True)
    o = oracle.invmgmt_episode(P, policy="base_stock", seed=4000)
    assert o["D"].tolist() == [19, 16, 16, 27, 25, 17, 16, 11, 19, 26, 20, 31, 24, 22, 19, 19, 16, 15, 23, 19, 20, 16,
                               15, 26, 23, 12, 19, 11, 22, 25]
    assert abs(seq_sum(o["reward"]) - 3891.9482044716797) < 1e-11
    assert o["I"][30].tolist() == [0, 150, 0] and o["B"][30].tolist() == [7, 0, 0, 0]
    P = pkg.InvManagementParams(backlog=False)
    o = oracle.invmgmt_episode(P, policy="base_stock", seed=4000)
    assert abs(seq_sum(o["reward"]) - 3704.961267089707) < 1e-11
    assert o["LS"].sum(axis=0).tolist() == [43, 0, 0, 0]
    assert oracle.invmgmt_episode(pkg.InvManagementParams(), actions=np.zeros((30, 3)), seed=42)["obs"][0].tolist() == \
        [100, 150, 200] + [0] * 30
