"""Device demand samplers vs scipy: chi-square goodness of fit, KS on the randomised PIT, moments, stream
independence (north-star: 'the on-device demand sampler must pass chi-square/KS agreement with scipy's Poisson')."""
import numpy as np
import pytest
from scipy import stats

import or_gym_inventory_b200 as pkg

pytestmark = pytest.mark.gpu
P_MIN = 1e-4   # each test draws millions of samples; a correct sampler fails this with probability 1e-4


def chi_square(samples, dist):
    samples = np.asarray(samples).ravel()
    n = samples.size
    lo, hi = int(samples.min()), int(samples.max())
    ks = np.arange(lo, hi + 1)
    obs = np.bincount(samples - lo, minlength=ks.size).astype(np.float64)
    exp = dist.pmf(ks) * n
    exp[0] += dist.cdf(lo - 1) * n          # fold the tails outside the observed range into the end bins
    exp[-1] += dist.sf(hi) * n
    # pool bins with small expectation
    o2, e2, ao, ae = [], [], 0.0, 0.0
    for o, e in zip(obs, exp):
        ao += o
        ae += e
        if ae >= 10:
            o2.append(ao); e2.append(ae); ao = ae = 0.0
    if ae > 0:
        if e2:
            o2[-1] += ao; e2[-1] += ae
        else:
            o2.append(ao); e2.append(ae)
    o2, e2 = np.array(o2), np.array(e2)
    if len(o2) < 2:
        return 1.0
    chi = ((o2 - e2) ** 2 / e2).sum()
    return stats.chi2.sf(chi, len(o2) - 1)


def ks_randomised(samples, dist, seed=0):
    """KS test for a discrete law through the randomised probability integral transform (uniform under H0)."""
    x = np.asarray(samples).ravel()
    v = np.random.default_rng(seed).random(x.size)
    u = dist.cdf(x - 1) + v * dist.pmf(x)
    return stats.kstest(u, "uniform").pvalue


@pytest.mark.parametrize("mu", [0.5, 5.0, 9.99, 10.0, 20.0, 200.0])
def test_alias_poisson_matches_scipy(mu):
    x = pkg.sample_demand("poisson", mu, seed=123, num_envs=100_000, per_env=30).cpu().numpy()
    d = stats.poisson(mu)
    assert chi_square(x, d) > P_MIN
    assert ks_randomised(x, d) > P_MIN
    assert abs(x.mean() - mu) < 5 * np.sqrt(mu / x.size) + 1e-12
    assert abs(x.var() - mu) < 0.01 * mu + 1e-3


def test_alias_poisson_mu_zero():
    assert (pkg.sample_demand("poisson", 0.0, seed=1, num_envs=1000, per_env=4) == 0).all()


@pytest.mark.parametrize("kind,p0,p1,dist", [
    ("binomial", 30, 0.3, stats.binom(30, 0.3)),
    ("randint", 2, 14, stats.randint(2, 15)),
    ("geometric", 0.12, 0.0, stats.geom(0.12)),
])
def test_alias_other_distributions(kind, p0, p1, dist):
    x = pkg.sample_demand(kind, p0, p1, seed=7, num_envs=50_000, per_env=30).cpu().numpy()
    assert chi_square(x, dist) > P_MIN
    assert ks_randomised(x, dist) > P_MIN


# table + inversion path: means below 260 (3.999 / 4.0 / 255.5 / 259.99 sit on table-grid edges);
# PTRS fallback: 260 and above
@pytest.mark.parametrize("mu,period", [(0.5, 3), (3.999, 2), (4.0, 3), (5.0, 3), (9.99, 2), (10.0, 3), (20.0, 3),
                                       (37.3, 4), (199.9, 3), (255.5, 6), (259.99, 3), (260.0, 3), (300.0, 2),
                                       (1000.5, 3)])
def test_per_env_mean_poisson_matches_scipy(mu, period):
    import torch
    m = torch.full((2_000_000,), mu, dtype=torch.float64, device="cuda")
    x = pkg.sample_poisson_mu(m, seed=2000, period=period).cpu().numpy()
    d = stats.poisson(mu)
    assert chi_square(x, d) > P_MIN
    assert ks_randomised(x, d) > P_MIN


def test_per_env_mean_poisson_mixture():
    """Newsvendor-like batch: mu_i uniform on [0, 200); check the PIT of every draw against its own mean."""
    import torch
    rng = np.random.default_rng(0)
    mu = rng.random(2_000_000) * 200.0
    x = pkg.sample_poisson_mu(torch.from_numpy(mu).cuda(), seed=5, period=0).cpu().numpy()
    v = rng.random(x.size)
    u = stats.poisson.cdf(x - 1, mu) + v * stats.poisson.pmf(x, mu)
    assert stats.kstest(u, "uniform").pvalue > P_MIN
    z = (x - mu)[mu > 1] / np.sqrt(mu[mu > 1])
    assert abs(z.mean()) < 5 / np.sqrt(z.size) and abs(z.var() - 1) < 0.01


def test_per_env_mean_poisson_degenerate_means():
    import torch
    m = torch.tensor([0.0, -1.0, float("nan"), 1e-300], dtype=torch.float64, device="cuda")
    assert (pkg.sample_poisson_mu(m, seed=1, period=0).cpu().numpy() == 0).all()


def test_per_env_mean_poisson_periods_are_independent():
    """Two consecutive periods share one Philox block (words x,y / z,w): their draws must be uncorrelated, and so
    must the two terms of the decomposition (alias word vs inversion word)."""
    import torch
    rng = np.random.default_rng(1)
    mu = rng.random(1_000_000) * 200.0
    m = torch.from_numpy(mu).cuda()
    a = pkg.sample_poisson_mu(m, seed=9, period=2).cpu().numpy()
    b = pkg.sample_poisson_mu(m, seed=9, period=3).cpu().numpy()
    c = pkg.sample_poisson_mu(m, seed=9, period=4).cpu().numpy()
    keep = mu > 1
    za, zb, zc = ((x - mu)[keep] / np.sqrt(mu[keep]) for x in (a, b, c))
    assert abs(np.corrcoef(za, zb)[0, 1]) < 5e-3 and abs(np.corrcoef(zb, zc)[0, 1]) < 5e-3
    # neighbouring envs (key = seed + id)
    assert abs(np.corrcoef(za[:-1], za[1:])[0, 1]) < 5e-3


def test_newsvendor_env_demand_large_mu_max():
    """mu_max = 1500 makes the env build its own (larger) table set; demand reported by step() must follow
    Poisson(mu_i) for every instance, and fixed means beyond the tables take the rejection path."""
    import torch
    N = 400_000
    env = pkg.NewsvendorEnv(num_envs=N, device="cuda:0", mu_max=1500, step_limit=4)
    env.reset(seed=11)
    mu = env.export_params().cpu().numpy()[:, 4]
    assert mu.max() > 1400
    z = torch.zeros(N, dtype=torch.float32, device="cuda")
    rng = np.random.default_rng(2)
    for t in range(3):
        x = env.step(z)[4]["demand"].cpu().numpy()
        u = stats.poisson.cdf(x - 1, mu) + rng.random(N) * stats.poisson.pmf(x, mu)
        assert stats.kstest(u, "uniform").pvalue > P_MIN
    env.close()
    env = pkg.NewsvendorEnv(num_envs=N, device="cuda:0", step_limit=4)
    fixed = np.tile(np.array([10.0, 5.0, 1.0, 2.0, 0.0]), (N, 1))
    fixed[:, 4] = rng.random(N) * 5000.0            # default tables end at 260
    env.reset(seed=12, options={"fixed_params": fixed})
    x = env.step(z)[4]["demand"].cpu().numpy()
    u = stats.poisson.cdf(x - 1, fixed[:, 4]) + rng.random(N) * stats.poisson.pmf(x, fixed[:, 4])
    assert stats.kstest(u, "uniform").pvalue > P_MIN
    env.close()


def test_stream_independence_and_keying():
    a = pkg.sample_demand("poisson", 20.0, seed=1, num_envs=200_000, per_env=30).cpu().numpy().astype(np.float64)
    # serial correlation across periods of one env and across neighbouring envs
    assert abs(np.corrcoef(a[:, :-1].ravel(), a[:, 1:].ravel())[0, 1]) < 5e-3
    assert abs(np.corrcoef(a[:-1].ravel(), a[1:].ravel())[0, 1]) < 5e-3
    # deterministic in (seed, env, period); env_offset shifts the global id; seed + 1 == next env id (key = seed + id)
    b = pkg.sample_demand("poisson", 20.0, seed=1, num_envs=1000, per_env=30, env_offset=500).cpu().numpy()
    assert np.array_equal(b, a[500:1500])
    c = pkg.sample_demand("poisson", 20.0, seed=99, num_envs=1000, per_env=30).cpu().numpy()
    assert not np.array_equal(c, a[:1000])
    # batch-size independence
    d = pkg.sample_demand("poisson", 20.0, seed=1, num_envs=7, per_env=30).cpu().numpy()
    assert np.array_equal(d, a[:7])
