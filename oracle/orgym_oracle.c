/*
 * orgym_oracle.c -- CPU restatement of the reference's step dynamics.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported, linked or
 * executed by the product package (or-gym-inventory_b200/); only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and there only as the checker / the timed CPU baseline.
 *
 * Parity status: PINNED by executing the reference.  The reference ships no
 * tests or golden vectors; the fixtures in tests/golden/ (npz files)
 * were produced by running the unmodified reference in the build container
 * (oracle/make_golden.py) and tests/test_oracle_golden.py checks every
 * function below against them bit for bit (state, observations, rewards) and,
 * for the sampled-demand path, against the reference's own PCG64 demand draws.
 *
 * Each function cites the reference file:line (paths relative to the
 * reference tree) or, for the third-party arithmetic on the path, the numpy
 * algorithm it restates (numpy is unpinned in requirements.txt:2; the oracle is
 * pinned to the container's numpy 2.3.5 behaviour).
 *
 * Build: see oracle/Makefile  (gcc -O2 -fPIC -shared -ffp-contract=off -pthread).
 * -ffp-contract=off matters: the reference never fuses a*b+c.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/orgym_b200.h"

#define ORC_API __attribute__((visibility("default")))

/* ======================================================================== *
 * numpy arithmetic helpers
 * ======================================================================== */

/* np.add.reduce over a contiguous 1-D float64 array: numpy's pairwise summation
 * (numpy/_core/src/umath/loops_utils.h.src, @TYPE@_pairwise_sum): plain loop for n < 8, eight running
 * partial sums for 8 <= n <= 128, recursive halves above. */
static double np_sum_f64(const double* a, int64_t n) {
    if (n < 8) {
        double res = 0.0;
        for (int64_t i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        int64_t i;
        for (i = 0; i < 8; i++) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return np_sum_f64(a, n2) + np_sum_f64(a + n2, n - n2);
    }
}

static float np_sum_f32(const float* a, int64_t n) {
    if (n < 8) {
        float res = 0.0f;
        for (int64_t i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        float r[8];
        int64_t i;
        for (i = 0; i < 8; i++) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return np_sum_f32(a, n2) + np_sum_f32(a + n2, n - n2);
    }
}

/* ======================================================================== *
 * numpy.random.Generator(PCG64(SeedSequence(seed)))  -- gymnasium's np_random
 * (test.py:5, used at inventory_management.py:197, newsvendor.py:102, network_management.py:303)
 * ======================================================================== */
typedef unsigned __int128 u128;
typedef struct {
    u128 state, inc;
} orc_rng_t;

/* numpy/random/bit_generator.pyx: SeedSequence.mix_entropy / generate_state for a single non-negative int seed */
static void seedseq_words(uint64_t seed, uint32_t* out, int n_words) {
    const uint32_t INIT_A = 0x43b0d7e5u, MULT_A = 0x931e8875u, INIT_B = 0x8b51f9ddu, MULT_B = 0x58f38dedu;
    const uint32_t MIX_L = 0xca01f9ddu, MIX_R = 0x4973f715u;
    uint32_t entropy[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    int n_ent = entropy[1] ? 2 : 1; /* int -> minimal little-endian uint32 array (0 -> [0]) */
    uint32_t pool[4], hc = INIT_A;
#define HASHMIX(v, dst)            \
    do {                           \
        uint32_t _x = (v) ^ hc;    \
        hc *= MULT_A;              \
        _x *= hc;                  \
        _x ^= _x >> 16;            \
        (dst) = _x;                \
    } while (0)
#define MIX(x, y, dst)                          \
    do {                                        \
        uint32_t _r = MIX_L * (x) - MIX_R * (y); \
        _r ^= _r >> 16;                         \
        (dst) = _r;                             \
    } while (0)
    for (int i = 0; i < 4; i++) HASHMIX(i < n_ent ? entropy[i] : 0u, pool[i]);
    for (int s = 0; s < 4; s++)
        for (int d = 0; d < 4; d++)
            if (s != d) {
                uint32_t h;
                HASHMIX(pool[s], h);
                MIX(pool[d], h, pool[d]);
            }
#undef HASHMIX
#undef MIX
    hc = INIT_B;
    for (int i = 0; i < n_words; i++) {
        uint32_t v = pool[i % 4] ^ hc;
        hc *= MULT_B;
        v *= hc;
        v ^= v >> 16;
        out[i] = v;
    }
}

#define PCG_MULT ((((u128)0x2360ED051FC65DA4ULL) << 64) | (u128)0x4385DF649FCCF645ULL)

/* numpy/random/src/pcg64/pcg64.h: pcg_setseq_128_srandom_r with (initstate, initseq) from SeedSequence */
static void orc_rng_seed(orc_rng_t* r, uint64_t seed) {
    uint32_t w[8];
    seedseq_words(seed, w, 8);
    uint64_t v[4];
    for (int i = 0; i < 4; i++) v[i] = (uint64_t)w[2 * i] | ((uint64_t)w[2 * i + 1] << 32);
    u128 initstate = ((u128)v[0] << 64) | v[1], initseq = ((u128)v[2] << 64) | v[3];
    r->state = 0;
    r->inc = (initseq << 1) | 1;
    r->state = r->state * PCG_MULT + r->inc;
    r->state += initstate;
    r->state = r->state * PCG_MULT + r->inc;
}

static inline uint64_t orc_rng_u64(orc_rng_t* r) {
    r->state = r->state * PCG_MULT + r->inc;
    uint64_t hi = (uint64_t)(r->state >> 64), lo = (uint64_t)r->state, x = hi ^ lo;
    unsigned rot = (unsigned)(r->state >> 122);
    return (x >> rot) | (x << ((-rot) & 63));
}

/* Generator.random(): (u64 >> 11) * 2**-53 */
static inline double orc_rng_double(orc_rng_t* r) { return (double)(orc_rng_u64(r) >> 11) * (1.0 / 9007199254740992.0); }

/* numpy/random/src/distributions/distributions.c: random_loggam */
static double np_loggam(double x) {
    static const double a[10] = {8.333333333333333e-02, -2.777777777777778e-03, 7.936507936507937e-04,
                                 -5.952380952380952e-04, 8.417508417508418e-04, -1.917526917526918e-03,
                                 6.410256410256410e-03,  -2.955065359477124e-02, 1.796443723688307e-01,
                                 -1.39243221690590e+00};
    double x0, x2, gl, gl0;
    long n = 0;
    if (x == 1.0 || x == 2.0) return 0.0;
    if (x < 7.0) n = (long)(7 - x);
    x0 = x + n;
    x2 = (1.0 / x0) * (1.0 / x0);
    gl0 = a[9];
    for (int k = 8; k >= 0; k--) {
        gl0 *= x2;
        gl0 += a[k];
    }
    gl = gl0 / x0 + 0.5 * 1.8378770664093453e+00 + (x0 - 0.5) * log(x0) - x0;
    if (x < 7.0)
        for (long k = 1; k <= n; k++) {
            gl -= log(x0 - 1.0);
            x0 -= 1.0;
        }
    return gl;
}

/* distributions.c: random_poisson = PTRS (Hoermann 1993) for lam >= 10, multiplication method below */
static int64_t orc_rng_poisson(orc_rng_t* r, double lam) {
    if (lam >= 10) {
        double slam = sqrt(lam), loglam = log(lam), b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
        double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2);
        for (;;) {
            double U = orc_rng_double(r) - 0.5, V = orc_rng_double(r), us = 0.5 - fabs(U);
            int64_t k = (int64_t)floor((2 * a / us + b) * U + lam + 0.43);
            if (us >= 0.07 && V <= vr) return k;
            if (k < 0 || (us < 0.013 && V > us)) continue;
            if ((log(V) + log(invalpha) - log(a / (us * us) + b)) <= (-lam + k * loglam - np_loggam((double)k + 1)))
                return k;
        }
    } else if (lam == 0) {
        return 0;
    } else {
        double enlam = exp(-lam), prod = 1.0;
        int64_t X = 0;
        for (;;) {
            prod *= orc_rng_double(r);
            if (prod > enlam)
                X += 1;
            else
                return X;
        }
    }
}

ORC_API void orc_np_poisson(uint64_t seed, double lam, int64_t n, int64_t* out) {
    orc_rng_t r;
    orc_rng_seed(&r, seed);
    for (int64_t i = 0; i < n; i++) out[i] = orc_rng_poisson(&r, lam);
}
ORC_API void orc_np_random(uint64_t seed, int64_t n, double* out) {
    orc_rng_t r;
    orc_rng_seed(&r, seed);
    for (int64_t i = 0; i < n; i++) out[i] = orc_rng_double(&r);
}

/* ======================================================================== *
 * Poisson quantile: scipy.stats.poisson.ppf(q, mu) = smallest integer k with cdf(k) >= q
 * (used by the classic-newsvendor and (s,S) drivers: benchmark_newsvendor.py:150,
 *  benchmark_newsvendor_sb3_rllib.py:366).  Independent implementation: exact pmf recurrence summed in
 * ascending order from 9 sigma below the mean; pinned against scipy in tests/test_oracle_golden.py.
 * ======================================================================== */
ORC_API double orc_poisson_ppf(double q, double mu) {
    if (!(q > 0.0)) return -1.0; /* scipy: ppf(0) = a - 1 */
    if (q >= 1.0) return INFINITY;
    double lo = floor(mu - 9.0 * sqrt(mu) - 9.0); /* mass below: < 3e-18 */
    if (lo < 0) lo = 0;
    double term = exp(-mu + lo * log(mu) - lgamma(lo + 1.0)), cdf = 0.0, k = lo;
    for (;;) {
        cdf += term;
        if (cdf >= q) return k;
        k += 1.0;
        term *= mu / k;
        if (term == 0.0 && k > mu) return k; /* q within rounding of 1 */
    }
}

/* ======================================================================== *
 * Serial multi-echelon env -- inventory_management.py
 * ======================================================================== */
#define MAXS (ORGYM_INV_MAX_STAGES + 1)

typedef struct {
    int m, n, T, backlog, lt_max, obs_dim;
    int64_t I0[MAXS], c[MAXS], L[MAXS];
    double up[MAXS], uc[MAXS], kc[MAXS], hc[MAXS], alpha;
    /* full histories like the reference (inventory_management.py:203-211) */
    int64_t *I, *B, *R, *S, *LS, *alog, *D;
    int t;
} inv_env_t;

static int inv_init(inv_env_t* e, const orgym_invmgmt_config_t* cfg) {
    memset(e, 0, sizeof(*e));
    e->m = cfg->num_stages;
    e->n = e->m - 1;
    e->T = cfg->periods;
    e->backlog = cfg->backlog;
    e->alpha = cfg->alpha;
    if (e->n < 1 || e->n > ORGYM_INV_MAX_STAGES || e->T < 1) return -1;
    for (int i = 0; i < e->n; i++) {
        e->I0[i] = cfg->init_inv[i];
        e->c[i] = cfg->capacity[i];
        e->L[i] = cfg->lead_time[i];
        if (e->L[i] > e->lt_max) e->lt_max = (int)e->L[i]; /* :100 */
    }
    for (int j = 0; j < e->m; j++) {
        e->up[j] = cfg->unit_price[j];
        e->uc[j] = cfg->unit_cost[j];
        e->kc[j] = cfg->demand_cost[j];
        e->hc[j] = cfg->holding_cost[j];
    }
    e->obs_dim = e->n * (e->lt_max + 1); /* :119 */
    int T = e->T, n = e->n, m = e->m;
    e->I = calloc((size_t)(T + 1) * n, 8);
    e->B = calloc((size_t)(T + 1) * m, 8);
    e->R = calloc((size_t)T * n, 8);
    e->S = calloc((size_t)T * m, 8);
    e->LS = calloc((size_t)T * m, 8);
    e->alog = calloc((size_t)T * n, 8);
    e->D = calloc((size_t)T, 8);
    return 0;
}
static void inv_free(inv_env_t* e) {
    free(e->I); free(e->B); free(e->R); free(e->S); free(e->LS); free(e->alog); free(e->D);
}

/* reset: inventory_management.py:186-222 */
static void inv_reset(inv_env_t* e) {
    int T = e->T, n = e->n, m = e->m;
    memset(e->I, 0, (size_t)(T + 1) * n * 8);
    memset(e->B, 0, (size_t)(T + 1) * m * 8);
    memset(e->R, 0, (size_t)T * n * 8);
    memset(e->S, 0, (size_t)T * m * 8);
    memset(e->LS, 0, (size_t)T * m * 8);
    memset(e->alog, 0, (size_t)T * n * 8);
    memset(e->D, 0, (size_t)T * 8);
    for (int i = 0; i < n; i++) e->I[i] = e->I0[i]; /* :215 */
    e->t = 0;
}

/* _get_obs: inventory_management.py:354-391 (oldest-first, left-aligned window; the docstring's order is wrong) */
static void inv_obs(const inv_env_t* e, int64_t* obs) {
    int n = e->n, t = e->t;
    memset(obs, 0, (size_t)e->obs_dim * 8);
    for (int i = 0; i < n; i++) obs[i] = e->I[(size_t)t * n + i]; /* :373 */
    if (t > 0) {
        int k = t < e->lt_max ? t : e->lt_max; /* :378 */
        memcpy(obs + n, e->alog + (size_t)(t - k) * n, (size_t)k * n * 8); /* :380-383 */
    }
}

/* step: inventory_management.py:224-352.  `action` as float64 (int64 actions are exact below 2**53).
 * Returns the discounted reward; *profit_out = undiscounted period profit. */
static double inv_step(inv_env_t* e, const double* action, int64_t demand_sample, double* profit_out) {
    int n = e->n, m = e->m, t = e->t;
    int64_t req[MAXS], cur[MAXS], Rf[MAXS], Ic[MAXS] = {0}, Sc[MAXS], Uc[MAXS];
    const int64_t* It = e->I + (size_t)t * n;
    const int64_t* Bt = e->B + (size_t)t * m;
    for (int i = 0; i < n; i++) {
        double a = action[i] > 0 ? action[i] : 0.0; /* np.maximum(action, 0) :250 */
        req[i] = (int64_t)a;                       /* .astype(int64) truncates */
        cur[i] = req[i] + (t >= 1 ? Bt[i + 1] : 0); /* :253-255 */
        int64_t r = cur[i] < e->c[i] ? cur[i] : e->c[i]; /* :263 */
        if (i < n - 1 && It[i + 1] < r) r = It[i + 1]; /* :260,265 supplier on-hand (may be negative); last stage: inf */
        Rf[i] = r;
        e->R[(size_t)t * n + i] = r;        /* :267 */
        e->alog[(size_t)t * n + i] = req[i]; /* :268 */
    }
    for (int i = 0; i < n; i++) { /* receipt :271-277 */
        Ic[i] = It[i];
        if (t - e->L[i] >= 0) Ic[i] += e->R[(size_t)(t - e->L[i]) * n + i];
    }
    int64_t d = demand_sample > 0 ? demand_sample : 0; /* :280 */
    e->D[t] = d;
    int64_t fill = d + (t >= 1 ? Bt[0] : 0); /* :284-286 */
    int64_t s0 = Ic[0] < fill ? Ic[0] : fill; /* :288 */
    Ic[0] -= s0;
    Sc[0] = s0;
    for (int i = 0; i < n; i++) Sc[i + 1] = Rf[i]; /* :294-295 */
    memcpy(e->S + (size_t)t * m, Sc, (size_t)m * 8);
    for (int i = 1; i < n; i++) Ic[i] -= Rf[i]; /* :300 (subtracts the stage's OWN inbound order) */
    Uc[0] = fill - s0;
    for (int i = 0; i < n; i++) Uc[i + 1] = cur[i] - Rf[i]; /* :303-304 */
    if (e->backlog)
        memcpy(e->B + (size_t)(t + 1) * m, Uc, (size_t)m * 8); /* :308 */
    else
        memcpy(e->LS + (size_t)t * m, Uc, (size_t)m * 8); /* :311 */
    double terms[MAXS];
    for (int j = 0; j < m; j++) { /* :315-321, elementwise float64 then np.sum */
        int64_t inv = j < n ? (Ic[j] > 0 ? Ic[j] : 0) : 0;
        double rev = e->up[j] * (double)Sc[j], pc = e->uc[j] * (double)Sc[j];
        double hold = e->hc[j] * (double)inv, pen = e->kc[j] * (double)Uc[j];
        terms[j] = ((rev - pc) - hold) - pen;
    }
    double profit = np_sum_f64(terms, m);
    double reward = pow(e->alpha, (double)t) * profit; /* :322  Python float ** int -> libm pow */
    memcpy(e->I + (size_t)(t + 1) * n, Ic, (size_t)n * 8); /* :326 */
    e->t = t + 1;
    if (profit_out) *profit_out = profit;
    return reward;
}

/* base-stock driver: benchmark_InvManagementBacklogEnv.py:152-198 (reads REQUESTED actions from action_log) */
static void inv_base_stock(const inv_env_t* e, double sf, double mu, double* action) {
    int n = e->n, t = e->t;
    for (int i = 0; i < n; i++) {
        int64_t pos = e->I[(size_t)t * n + i];
        int64_t L = e->L[i];
        if (L > 0 && t > 0) {
            int64_t lo = t - L > 0 ? t - L : 0;
            for (int64_t p = lo; p < t; p++) pos += e->alog[(size_t)p * n + i];
        }
        double target = ((double)(L + 1) * mu) * sf; /* (lead_times + 1) * mu * safety_factor (:186), float64 */
        double q = target - (double)pos;
        if (!(q > 0)) q = 0;                     /* np.maximum(0, .) */
        if (q > (double)e->c[i]) q = (double)e->c[i]; /* np.clip(., low=0, high=c) */
        action[i] = (double)(int64_t)q;          /* .astype(int64) */
    }
}

/* Replay / drive one episode.
 * policy: 0 = actions given [T][n] float64, 1 = base-stock (pparam[0]=sf, pparam[1]=mu).
 * demand: int64[T] or NULL -> sample Poisson(mu_demand) from Generator(PCG64(SeedSequence(seed))) like the reference.
 * Outputs (any may be NULL): obs[T+1][obs_dim], reward[T], profit[T], I[T+1][n], B[T+1][m], S[T][m], LS[T][m],
 * R[T][n], alog[T][n], D[T]. */
ORC_API int orc_invmgmt_episode(const orgym_invmgmt_config_t* cfg, int policy, const double* pparam,
                                const double* actions, const int64_t* demand, uint64_t seed, int64_t* obs,
                                double* reward, double* profit, int64_t* I, int64_t* B, int64_t* S, int64_t* LS,
                                int64_t* R, int64_t* alog, int64_t* D) {
    inv_env_t e;
    if (inv_init(&e, cfg)) return -1;
    inv_reset(&e);
    orc_rng_t rng;
    orc_rng_seed(&rng, seed);
    int T = e.T, n = e.n, m = e.m;
    if (obs) inv_obs(&e, obs);
    for (int t = 0; t < T; t++) {
        double a[MAXS];
        if (policy == 1)
            inv_base_stock(&e, pparam[0], pparam[1], a);
        else
            memcpy(a, actions + (size_t)t * n, (size_t)n * 8);
        int64_t d;
        if (demand)
            d = demand[t];
        else if (cfg->dist.kind == ORGYM_DIST_POISSON)
            d = orc_rng_poisson(&rng, cfg->dist.p0); /* :172 */
        else if (cfg->dist.kind == ORGYM_DIST_USER)
            d = t < cfg->dist.user_D_len ? cfg->dist.user_D[t] : 0; /* :182 */
        else {
            inv_free(&e);
            return -3;
        }
        double pr;
        double r = inv_step(&e, a, d, &pr);
        if (reward) reward[t] = r;
        if (profit) profit[t] = pr;
        if (obs) inv_obs(&e, obs + (size_t)(t + 1) * e.obs_dim);
    }
    if (I) memcpy(I, e.I, (size_t)(T + 1) * n * 8);
    if (B) memcpy(B, e.B, (size_t)(T + 1) * m * 8);
    if (S) memcpy(S, e.S, (size_t)T * m * 8);
    if (LS) memcpy(LS, e.LS, (size_t)T * m * 8);
    if (R) memcpy(R, e.R, (size_t)T * n * 8);
    if (alog) memcpy(alog, e.alog, (size_t)T * n * 8);
    if (D) memcpy(D, e.D, (size_t)T * 8);
    inv_free(&e);
    return 0;
}

/* ---- CPU baseline: the reference's evaluation loop (benchmark_InvManagementLostSalesEnv.py:239-302:
 * reset(seed) ; while not truncated: a = agent(obs); step(a)) for `episodes` episodes split over `threads`
 * host threads; each step also assembles the observation like the reference does.  Returns total env-steps;
 * *ret_sum gets the sum of episode returns (keeps the work observable). */
typedef struct {
    const orgym_invmgmt_config_t* cfg;
    int policy;
    double pparam[4];
    uint64_t seed0;
    int64_t ep_begin, ep_end, steps;
    double ret_sum;
} inv_job_t;

static void* inv_worker(void* arg) {
    inv_job_t* j = (inv_job_t*)arg;
    inv_env_t e;
    if (inv_init(&e, j->cfg)) return NULL;
    int64_t* obs = malloc((size_t)e.obs_dim * 8);
    orc_rng_t rng, arng;
    double mu = j->cfg->dist.p0;
    for (int64_t ep = j->ep_begin; ep < j->ep_end; ep++) {
        inv_reset(&e);
        orc_rng_seed(&rng, j->seed0 + (uint64_t)ep);
        orc_rng_seed(&arng, (j->seed0 + (uint64_t)ep) ^ 0x9E3779B97F4A7C15ULL);
        inv_obs(&e, obs);
        double ret = 0;
        for (int t = 0; t < e.T; t++) {
            double a[MAXS];
            if (j->policy == 1)
                inv_base_stock(&e, j->pparam[0], j->pparam[1], a);
            else
                for (int i = 0; i < e.n; i++) a[i] = (double)(orc_rng_u64(&arng) % (uint64_t)(e.c[i] + 1));
            ret += inv_step(&e, a, orc_rng_poisson(&rng, mu), NULL);
            inv_obs(&e, obs);
        }
        j->ret_sum += ret + (double)obs[0] * 0.0;
        j->steps += e.T;
    }
    free(obs);
    inv_free(&e);
    return NULL;
}

ORC_API int64_t orc_invmgmt_bench(const orgym_invmgmt_config_t* cfg, int policy, const double* pparam,
                                  uint64_t seed0, int64_t episodes, int threads, double* ret_sum) {
    if (threads < 1) threads = 1;
    pthread_t* th = malloc(sizeof(pthread_t) * (size_t)threads);
    inv_job_t* jobs = calloc((size_t)threads, sizeof(inv_job_t));
    for (int k = 0; k < threads; k++) {
        jobs[k].cfg = cfg;
        jobs[k].policy = policy;
        if (pparam) memcpy(jobs[k].pparam, pparam, sizeof(double) * 4);
        jobs[k].seed0 = seed0;
        jobs[k].ep_begin = episodes * k / threads;
        jobs[k].ep_end = episodes * (k + 1) / threads;
        pthread_create(&th[k], NULL, inv_worker, &jobs[k]);
    }
    int64_t steps = 0;
    double rs = 0;
    for (int k = 0; k < threads; k++) {
        pthread_join(th[k], NULL);
        steps += jobs[k].steps;
        rs += jobs[k].ret_sum;
    }
    if (ret_sum) *ret_sum = rs;
    free(th);
    free(jobs);
    return steps;
}

/* ======================================================================== *
 * Newsvendor -- newsvendor.py.  Scalars carry the numpy "kind" that decides the precision of each
 * binary operation under NumPy >= 2 (NEP 50) promotion: PY = Python int/float (weak), F32 = np.float32,
 * F64 = np.float64.  any F64 -> float64 op; else any F32 -> both operands rounded to float32, float32 op;
 * else Python float op.
 * ======================================================================== */
enum { K_PY = 0, K_F32 = 1, K_F64 = 2 };
typedef struct {
    double v;
    int k;
} sc_t;
static inline sc_t sc(double v, int k) {
    sc_t s = {v, k};
    return s;
}
static inline sc_t sc_bin(sc_t a, sc_t b, char op) {
    if (a.k == K_F64 || b.k == K_F64 || (a.k == K_PY && b.k == K_PY)) {
        double r = op == '*' ? a.v * b.v : a.v - b.v;
        return sc(r, (a.k == K_PY && b.k == K_PY) ? K_PY : K_F64);
    }
    float x = (float)a.v, y = (float)b.v;
    float r = op == '*' ? x * y : x - y;
    return sc((double)r, K_F32);
}

typedef struct {
    int L, T, obs_dim;
    double max_inv, max_q, p_max, h_max, k_max, mu_max;
    double price, cost, h, k, mu; /* Python floats (newsvendor.py:105-111) */
    float state[5 + ORGYM_NV_MAX_LEAD];
    int step_count;
} nv_env_t;

static void nv_init(nv_env_t* e, const orgym_newsvendor_config_t* c) {
    memset(e, 0, sizeof(*e));
    e->L = c->lead_time;
    e->T = c->step_limit;
    e->obs_dim = e->L + 5; /* :76 */
    e->max_inv = c->max_inventory;
    e->max_q = c->max_order_quantity;
    e->p_max = c->p_max; e->h_max = c->h_max; e->k_max = c->k_max; e->mu_max = c->mu_max;
}

/* reset: newsvendor.py:100-123; u[5] are the five Generator.random() draws (or NULL with fixed params) */
static void nv_reset(nv_env_t* e, const double* u, const double* fixed) {
    if (fixed) {
        e->price = fixed[0]; e->cost = fixed[1]; e->h = fixed[2]; e->k = fixed[3]; e->mu = fixed[4];
    } else {
        double x = u[0] * e->p_max;
        e->price = x > 1 ? x : 1; /* max(1, .) :105 */
        x = u[1] * e->price;
        e->cost = x > 1 ? x : 1; /* :107 */
        e->h = u[2] * (e->h_max < e->cost ? e->h_max : e->cost); /* min(cost, h_max) :109 */
        e->k = u[3] * e->k_max;
        e->mu = u[4] * e->mu_max;
    }
    memset(e->state, 0, sizeof(e->state));
    e->state[0] = (float)e->price; e->state[1] = (float)e->cost; e->state[2] = (float)e->h;
    e->state[3] = (float)e->k; e->state[4] = (float)e->mu; /* :115 */
    e->step_count = 0;
}

/* step: newsvendor.py:125-204.  action = the float32 element of the size-1 action array. parts[4] optional. */
static double nv_step(nv_env_t* e, float action, int64_t demand, double* parts) {
    e->step_count += 1;
    double q = (double)action; /* .item() -> Python float :131 */
    q = q < 0 ? 0 : (q > e->max_q ? e->max_q : q); /* np.clip -> np.float64 :132 */
    sc_t order = sc(q, K_F64);
    float* pipe = e->state + 5;
    float s = np_sum_f32(pipe, e->L);                      /* :135 */
    sc_t inv = e->L > 0 ? sc((double)pipe[0], K_F32) : order; /* :136-139 */
    sc_t cap = sc((double)((float)e->max_inv - s), K_F32);  /* int - np.float32 -> float32 :143 */
    sc_t mn = cap.v < order.v ? cap : order;               /* min(order_qty, cap) */
    sc_t oq = mn.v > 0 ? mn : sc(0.0, K_PY);               /* max(0, .) */
    sc_t dem = sc((double)demand, K_PY);
    sc_t su = dem.v < inv.v ? dem : inv;                   /* min(inv_on_hand, demand) :149 */
    sc_t rev = sc_bin(su, sc(e->price, K_PY), '*');        /* :150 */
    sc_t ex = sc_bin(inv, dem, '-');
    sc_t exs = ex.v > 0 ? ex : sc(0.0, K_PY);              /* :152 */
    sc_t sh = sc_bin(dem, inv, '-');
    sc_t shs = sh.v > 0 ? sh : sc(0.0, K_PY);              /* :153 */
    sc_t pc = sc_bin(oq, sc(e->cost, K_PY), '*');          /* :162 */
    sc_t hc = sc_bin(exs, sc(e->h, K_PY), '*');            /* :166 */
    sc_t lp = sc_bin(shs, sc(e->k, K_PY), '*');            /* :167 */
    sc_t rew = sc_bin(sc_bin(sc_bin(rev, pc, '-'), hc, '-'), lp, '-'); /* :170 */
    if (e->L > 0) { /* :174-183 */
        for (int i = 0; i + 1 < e->L; i++) pipe[i] = pipe[i + 1];
        pipe[e->L - 1] = (float)oq.v;
    }
    if (parts) {
        parts[0] = rev.v; parts[1] = pc.v; parts[2] = hc.v; parts[3] = lp.v;
    }
    return rew.v;
}

/* drivers, computed on the float32 observation like the reference agents do */
static float nv_clip_f32(float q, float hi) { return q < 0 ? 0 : (q > hi ? hi : q); }

static float nv_policy(const nv_env_t* e, int policy, const double* pp) {
    const float* o = e->state;
    float h = o[2], k = o[3], mu = o[4];
    float pos = np_sum_f32(o + 5, e->L);
    float hi = (float)e->max_q;
    if (policy == ORGYM_NV_POLICY_ORDER_UP_TO) { /* benchmark_newsvendor.py:103-111, all float32 */
        float target = (mu * (float)(e->L + 1)) * (float)pp[0];
        float q = target - pos;
        return nv_clip_f32(q > 0 ? q : 0, hi);
    }
    if (policy == ORGYM_NV_POLICY_CLASSIC) { /* benchmark_newsvendor.py:113-161, cr_method k_vs_h */
        float hk = h + k;
        if (hk <= 1e-6f || k < 0 || h < 0) {
            float q = mu * (float)(e->L + 1) - pos;
            return nv_clip_f32(q > 0 ? q : 0, hi);
        }
        float cr = k / hk;
        float eff = (mu * (float)(e->L + 1)) * (float)pp[0];
        double m = eff > 1e-6f ? (double)eff : 1e-6; /* max(1e-6, eff_mu): float32 comparison */
        double lvl = orc_poisson_ppf((double)cr, m);
        double q = lvl - (double)pos; /* np.float64 - np.float32 -> float64 */
        if (!(q > 0)) q = 0;
        if (q > (double)hi) q = (double)hi;
        return (float)q;
    }
    if (policy == ORGYM_NV_POLICY_SS) { /* benchmark_newsvendor_sb3_rllib.py:363-371 */
        double s_lvl = 0;
        if (h + k > 1e-6f) {
            float cr = k / (h + k);
            cr = cr < 0.001f ? 0.001f : (cr > 0.999f ? 0.999f : cr); /* np.clip on float32 keeps float32 */
            float eff = mu * (float)(e->L + 1);
            double m = eff > 1e-6f ? (double)eff : 1e-6;
            s_lvl = orc_poisson_ppf((double)cr, m);
        }
        double s_level = s_lvl > 0 ? s_lvl : 0, S_level = s_level * pp[0];
        double q = 0;
        if ((double)pos < s_level) {
            q = S_level - (double)pos;
            if (!(q > 0)) q = 0;
        }
        if (q > (double)hi) q = (double)hi;
        return (float)q;
    }
    return 0;
}

/* One episode.  If params == NULL the five reset uniforms and the demand come from Generator(PCG64(SeedSequence(seed)))
 * exactly as in the reference; otherwise params[5] are used (fixed_params) and demand must be given or is sampled
 * after zero uniform draws.  demand: int64[T] or NULL.  actions: float32[T] (policy 0).
 * Outputs: params_out[5], obs[T+1][obs_dim] float32, reward[T], parts[T][4], demand_out[T], actions_out[T]. */
ORC_API int orc_newsvendor_episode(const orgym_newsvendor_config_t* cfg, int policy, const double* pparam,
                                   const float* actions, const int64_t* demand, const double* params, uint64_t seed,
                                   double* params_out, float* obs, double* reward, double* parts,
                                   int64_t* demand_out, float* actions_out) {
    nv_env_t e;
    nv_init(&e, cfg);
    if (e.L < 0 || e.L > ORGYM_NV_MAX_LEAD) return -1;
    orc_rng_t rng;
    orc_rng_seed(&rng, seed);
    if (params)
        nv_reset(&e, NULL, params);
    else {
        double u[5];
        /* the draws interleave with their use in the reference, but each is one Generator.random() in order */
        for (int i = 0; i < 5; i++) u[i] = orc_rng_double(&rng);
        nv_reset(&e, u, NULL);
    }
    if (params_out) {
        params_out[0] = e.price; params_out[1] = e.cost; params_out[2] = e.h; params_out[3] = e.k; params_out[4] = e.mu;
    }
    if (obs) memcpy(obs, e.state, (size_t)e.obs_dim * 4);
    for (int t = 0; t < e.T; t++) {
        float a = policy == 0 ? actions[t] : nv_policy(&e, policy, pparam);
        int64_t d = demand ? demand[t] : orc_rng_poisson(&rng, e.mu); /* :146 */
        double r = nv_step(&e, a, d, parts ? parts + 4 * (size_t)t : NULL);
        if (reward) reward[t] = r;
        if (demand_out) demand_out[t] = d;
        if (actions_out) actions_out[t] = a;
        if (obs) memcpy(obs + (size_t)(t + 1) * e.obs_dim, e.state, (size_t)e.obs_dim * 4);
    }
    return 0;
}

typedef struct {
    const orgym_newsvendor_config_t* cfg;
    int policy;
    double pparam[4];
    uint64_t seed0;
    int64_t ep_begin, ep_end, steps;
    double ret_sum;
} nv_job_t;

static void* nv_worker(void* arg) {
    nv_job_t* j = (nv_job_t*)arg;
    nv_env_t e;
    nv_init(&e, j->cfg);
    float obs[5 + ORGYM_NV_MAX_LEAD];
    for (int64_t ep = j->ep_begin; ep < j->ep_end; ep++) {
        orc_rng_t rng;
        orc_rng_seed(&rng, j->seed0 + (uint64_t)ep);
        double u[5];
        for (int i = 0; i < 5; i++) u[i] = orc_rng_double(&rng);
        nv_reset(&e, u, NULL);
        double ret = 0;
        for (int t = 0; t < e.T; t++) {
            float a = nv_policy(&e, j->policy, j->pparam);
            ret += nv_step(&e, a, orc_rng_poisson(&rng, e.mu), NULL);
            memcpy(obs, e.state, (size_t)e.obs_dim * 4); /* _get_obs copy :208 */
        }
        j->ret_sum += ret + obs[0] * 0.0;
        j->steps += e.T;
    }
    return NULL;
}

ORC_API int64_t orc_newsvendor_bench(const orgym_newsvendor_config_t* cfg, int policy, const double* pparam,
                                     uint64_t seed0, int64_t episodes, int threads, double* ret_sum) {
    if (threads < 1) threads = 1;
    pthread_t* th = malloc(sizeof(pthread_t) * (size_t)threads);
    nv_job_t* jobs = calloc((size_t)threads, sizeof(nv_job_t));
    for (int k = 0; k < threads; k++) {
        jobs[k].cfg = cfg;
        jobs[k].policy = policy;
        if (pparam) memcpy(jobs[k].pparam, pparam, sizeof(double) * 4);
        jobs[k].seed0 = seed0;
        jobs[k].ep_begin = episodes * k / threads;
        jobs[k].ep_end = episodes * (k + 1) / threads;
        pthread_create(&th[k], NULL, nv_worker, &jobs[k]);
    }
    int64_t steps = 0;
    double rs = 0;
    for (int k = 0; k < threads; k++) {
        pthread_join(th[k], NULL);
        steps += jobs[k].steps;
        rs += jobs[k].ret_sum;
    }
    if (ret_sum) *ret_sum = rs;
    free(th);
    free(jobs);
    return steps;
}

/* ======================================================================== *
 * Network env -- network_management.py (float64 state, Python evaluation order)
 * ======================================================================== */
typedef struct {
    const orgym_netinv_config_t* c;
    int J, E, M, T, obs_dim;
    /* histories (network_management.py:315-321) */
    double *X, *Y, *R, *S, *D, *U, *P;
    int t;
} net_env_t;

static int net_init(net_env_t* e, const orgym_netinv_config_t* c) {
    memset(e, 0, sizeof(*e));
    e->c = c;
    e->J = c->num_main; e->E = c->num_reorder; e->M = c->num_retail; e->T = c->num_periods;
    int sumL = 0;
    for (int i = 0; i < e->E; i++) sumL += c->re_lead[i];
    e->obs_dim = e->M + e->J + sumL; /* :190 */
    int T = e->T;
    e->X = calloc((size_t)(T + 1) * e->J + 1, 8);
    e->Y = calloc((size_t)(T + 1) * e->E + 1, 8);
    e->R = calloc((size_t)T * e->E + 1, 8);
    e->S = calloc((size_t)T * (e->E + e->M) + 1, 8);
    e->D = calloc((size_t)T * e->M + 1, 8);
    e->U = calloc((size_t)(T + 1) * e->M + 1, 8);
    e->P = calloc((size_t)T * e->J + 1, 8);
    return 0;
}
static void net_free(net_env_t* e) {
    free(e->X); free(e->Y); free(e->R); free(e->S); free(e->D); free(e->U); free(e->P);
}
static void net_reset(net_env_t* e) { /* :301-332 */
    int T = e->T;
    memset(e->X, 0, ((size_t)(T + 1) * e->J) * 8);
    memset(e->Y, 0, ((size_t)(T + 1) * e->E) * 8);
    memset(e->R, 0, ((size_t)T * e->E) * 8);
    memset(e->S, 0, ((size_t)T * (e->E + e->M)) * 8);
    memset(e->D, 0, ((size_t)T * e->M) * 8);
    memset(e->U, 0, ((size_t)(T + 1) * e->M) * 8);
    memset(e->P, 0, ((size_t)T * e->J) * 8);
    for (int j = 0; j < e->J; j++) e->X[j] = e->c->node_I0[j]; /* :326 */
    e->t = 0;
}
/* _get_obs: :334-413 */
static void net_obs(const net_env_t* e, float* obs) {
    int t = e->t, k = 0;
    for (int r = 0; r < e->M; r++) obs[k++] = (float)e->U[(size_t)t * e->M + r];
    for (int j = 0; j < e->J; j++) obs[k++] = (float)e->X[(size_t)t * e->J + j];
    for (int i = 0; i < e->E; i++) {
        int L = e->c->re_lead[i];
        if (L == 0) continue; /* :353 */
        int have = t < L ? t : L;
        for (int z = 0; z < L - have; z++) obs[k++] = 0.0f; /* left zero padding :375-378 */
        for (int p = t - have; p < t; p++) obs[k++] = (float)e->R[(size_t)p * e->E + i];
    }
}

/* Python round() of a numpy float32: round-half-to-even to an integer */
static double py_round_f32(float a) { return nearbyint((double)a); /* default FE_TONEAREST = ties-to-even */ }

/* step: :436-635.  demand[M] = this period's realised demand per retail link (already max(0,int(round(.)))) */
static double net_step(net_env_t* e, const float* action, const double* demand, double* profit_out) {
    const orgym_netinv_config_t* c = e->c;
    int t = e->t, J = e->J, E = e->E, M = e->M;
    double cons[ORGYM_NET_MAX_NODES], arr[ORGYM_NET_MAX_NODES], Xn[ORGYM_NET_MAX_NODES];
    const double* Xt = e->X + (size_t)t * J;
    double* Rt = e->R + (size_t)t * E;
    double* St = e->S + (size_t)t * (E + M);
    for (int j = 0; j < J; j++) cons[j] = 0.0;
    for (int i = 0; i < E; i++) { /* :448-490, sequential greedy allocation in sorted link order */
        double req = py_round_f32(action[i]);
        if (!(req > 0)) req = 0; /* max(0, round(a)) :449 */
        int s = c->re_supplier[i];
        double f = 0;
        if (s == -1)
            f = req; /* raw material :453-455 */
        else if (s >= 0) {
            double avail = Xt[s] - cons[s]; /* :459 */
            if (!(avail > 0)) avail = 0;    /* max(0, .) :460 */
            double oa = avail;
            if (c->node_is_factory[s]) { /* :464-478 */
                double mp = c->node_v[s] * avail;
                double lim = mp < c->node_C[s] ? mp : c->node_C[s]; /* min(C, v*avail) :476 */
                if (lim < oa) oa = lim;
            }
            f = oa < req ? oa : req; /* :481 */
            cons[s] += f / c->node_v[s]; /* :484-485 */
        }
        Rt[i] = f; /* :488 */
        St[i] = f; /* :490 */
    }
    double* Yn = e->Y + (size_t)(t + 1) * E;
    const double* Yt = e->Y + (size_t)t * E;
    for (int i = 0; i < E; i++) { /* :494-511 */
        int L = c->re_lead[i];
        double arriving = (t - L >= 0) ? e->R[(size_t)(t - L) * E + i] : 0.0;
        Yn[i] = (Yt[i] - arriving) + Rt[i];
    }
    for (int j = 0; j < J; j++) { /* :516-523 arrivals in predecessor (adjacency) order */
        double a = 0.0;
        for (int z = c->pred_ptr[j]; z < c->pred_ptr[j + 1]; z++) {
            int i = c->pred_idx[z], L = c->re_lead[i];
            if (t - L >= 0) a += e->R[(size_t)(t - L) * E + i];
        }
        arr[j] = a;
        Xn[j] = (Xt[j] + arr[j]) - cons[j]; /* :528 */
    }
    const double* Ut = e->U + (size_t)t * M;
    double* Un = e->U + (size_t)(t + 1) * M;
    for (int r = 0; r < M; r++) { /* :536-566 */
        int j = c->rt_retailer[r];
        double d = demand[r];
        e->D[(size_t)t * M + r] = d;
        double fill = d + Ut[r];
        double invr = Xn[j] > 0 ? Xn[j] : 0; /* max(0, X) :547 */
        double s = invr < fill ? invr : fill; /* min(fill, inv) :548 */
        St[E + r] = s;
        Xn[j] -= s;
        Un[r] = c->backlog ? fill - s : 0.0; /* :560-563 (lost-sales amount is discarded) */
    }
    memcpy(e->X + (size_t)(t + 1) * J, Xn, (size_t)J * 8); /* :571 */
    double total = 0;
    for (int j = 0; j < J; j++) { /* :578-613, Python sum() order = adjacency order */
        double SR = 0, PC = 0, HCp = 0, sold = 0, UP = 0;
        for (int z = c->succ_ptr[j]; z < c->succ_ptr[j + 1]; z++) {
            int l = c->succ_idx[z];
            double p = l < E ? c->re_p[l] : c->rt_p[l - E];
            SR += p * St[l];  /* :582 */
            sold += St[l];    /* :599 */
            if (l >= E) UP += c->rt_b[l - E] * Un[l - E]; /* :608 */
        }
        for (int z = c->pred_ptr[j]; z < c->pred_ptr[j + 1]; z++) {
            int i = c->pred_idx[z];
            PC += c->re_p[i] * Rt[i]; /* :586 */
            HCp += c->re_g[i] * (Yn[i] > 0 ? Yn[i] : 0); /* :591 */
        }
        double HC = c->node_h[j] * (Xn[j] > 0 ? Xn[j] : 0) + HCp; /* :590,593 */
        double OC = 0;
        if (c->node_is_factory[j]) OC = c->node_v[j] > 0 ? c->node_o[j] * (sold / c->node_v[j]) : 0; /* :597-601 */
        if (!c->node_is_retail[j]) UP = 0; /* :605 */
        double P = (((SR - PC) - OC) - HC) - UP; /* :611 */
        e->P[(size_t)t * J + j] = P;
        total += P;
    }
    double reward = pow(c->alpha, (double)t) * total; /* :619 */
    e->t = t + 1;
    if (profit_out) *profit_out = total;
    return reward;
}

/* One episode.  policy 0: actions float32[T][E]; policy 1: constant action float32[E].
 * demand: float64[T][M] realised demands, or NULL -> for every retail link, in retail_links order, one Poisson draw
 * per period from Generator(PCG64(SeedSequence(seed))) (the reference's default samplers, :125-127, :263).
 * Outputs (may be NULL): obs[T+1][obs_dim], reward[T], profit[T], X[T+1][J], Y[T+1][E], U[T+1][M], R[T][E],
 * S[T][E+M] (reorder links then retail links), D[T][M], P[T][J]. */
ORC_API int orc_netinv_episode(const orgym_netinv_config_t* cfg, int policy, const float* actions,
                               const double* demand, uint64_t seed, float* obs, double* reward, double* profit,
                               double* X, double* Y, double* U, double* R, double* S, double* D, double* P) {
    net_env_t e;
    if (cfg->num_main > ORGYM_NET_MAX_NODES) return -1;
    net_init(&e, cfg);
    net_reset(&e);
    orc_rng_t rng;
    orc_rng_seed(&rng, seed);
    int T = e.T, E = e.E, M = e.M, J = e.J;
    if (obs) net_obs(&e, obs);
    double dem[ORGYM_NET_MAX_RETAIL];
    for (int t = 0; t < T; t++) {
        const float* a = policy == 1 ? actions : actions + (size_t)t * E;
        for (int r = 0; r < M; r++) {
            if (demand)
                dem[r] = demand[(size_t)t * M + r];
            else if (cfg->rt_dist[r].kind == ORGYM_DIST_POISSON) {
                double v = (double)orc_rng_poisson(&rng, cfg->rt_dist[r].p0);
                dem[r] = v > 0 ? v : 0;
            } else if (cfg->rt_dist[r].kind == ORGYM_DIST_USER) {
                int idx = t < cfg->rt_dist[r].user_D_len - 1 ? t : cfg->rt_dist[r].user_D_len - 1; /* :255 */
                dem[r] = (double)cfg->rt_dist[r].user_D[idx];
                if (dem[r] < 0) dem[r] = 0;
            } else {
                net_free(&e);
                return -3;
            }
        }
        double pr;
        double r = net_step(&e, a, dem, &pr);
        if (reward) reward[t] = r;
        if (profit) profit[t] = pr;
        if (obs) net_obs(&e, obs + (size_t)(t + 1) * e.obs_dim);
    }
    if (X) memcpy(X, e.X, (size_t)(T + 1) * J * 8);
    if (Y) memcpy(Y, e.Y, (size_t)(T + 1) * E * 8);
    if (U) memcpy(U, e.U, (size_t)(T + 1) * M * 8);
    if (R) memcpy(R, e.R, (size_t)T * E * 8);
    if (S) memcpy(S, e.S, (size_t)T * (E + M) * 8);
    if (D) memcpy(D, e.D, (size_t)T * M * 8);
    if (P) memcpy(P, e.P, (size_t)T * J * 8);
    net_free(&e);
    return 0;
}

typedef struct {
    const orgym_netinv_config_t* cfg;
    const float* action;
    uint64_t seed0;
    int64_t ep_begin, ep_end, steps;
    double ret_sum;
} net_job_t;

static void* net_worker(void* arg) {
    net_job_t* j = (net_job_t*)arg;
    net_env_t e;
    net_init(&e, j->cfg);
    float* obs = malloc((size_t)e.obs_dim * 4 + 4);
    double dem[ORGYM_NET_MAX_RETAIL];
    for (int64_t ep = j->ep_begin; ep < j->ep_end; ep++) {
        net_reset(&e);
        orc_rng_t rng;
        orc_rng_seed(&rng, j->seed0 + (uint64_t)ep);
        double ret = 0;
        for (int t = 0; t < e.T; t++) {
            for (int r = 0; r < e.M; r++) dem[r] = (double)orc_rng_poisson(&rng, j->cfg->rt_dist[r].p0);
            ret += net_step(&e, j->action, dem, NULL);
            net_obs(&e, obs);
        }
        j->ret_sum += ret + obs[0] * 0.0;
        j->steps += e.T;
    }
    free(obs);
    net_free(&e);
    return NULL;
}

ORC_API int64_t orc_netinv_bench(const orgym_netinv_config_t* cfg, const float* const_action, uint64_t seed0,
                                 int64_t episodes, int threads, double* ret_sum) {
    if (threads < 1) threads = 1;
    pthread_t* th = malloc(sizeof(pthread_t) * (size_t)threads);
    net_job_t* jobs = calloc((size_t)threads, sizeof(net_job_t));
    for (int k = 0; k < threads; k++) {
        jobs[k].cfg = cfg;
        jobs[k].action = const_action;
        jobs[k].seed0 = seed0;
        jobs[k].ep_begin = episodes * k / threads;
        jobs[k].ep_end = episodes * (k + 1) / threads;
        pthread_create(&th[k], NULL, net_worker, &jobs[k]);
    }
    int64_t steps = 0;
    double rs = 0;
    for (int k = 0; k < threads; k++) {
        pthread_join(th[k], NULL);
        steps += jobs[k].steps;
        rs += jobs[k].ret_sum;
    }
    if (ret_sum) *ret_sum = rs;
    free(th);
    free(jobs);
    return steps;
}

ORC_API int orc_version(void) { return ORGYM_VERSION; }
