// poisson_mu.cuh -- Poisson sampling with a per-instance mean (newsvendor demand, orgym_sample_poisson_mu) and the
// reciprocal / log-factorial helpers it shares with the Poisson quantile.  Kept out of common.cuh so that the serial
// and network env kernels (and the profile provenance hashes of bench.py) do not depend on it.
#pragma once
#include "common.cuh"

// exact reciprocals 1/k for 1 <= k <= 2048 (16 KB, L1-resident): the Poisson recurrences p_k = p_{k-1} * mu / k use
// a table load and a multiply instead of a float64 division.  Filled once per device by orgym_init_tables().
#define ORGYM_RCP_N 2048
// device pointer to the table of `device` (allocated and filled on first use, kept for the life of the process)
int orgym_rcp_table(int device, const double** out);
__device__ __forceinline__ double rcp_int(const double* __restrict__ tab, double k) {
    return k <= (double)ORGYM_RCP_N ? tab[(int)k] : __drcp_rn(k);
}

// log(k!) for integer-valued k >= 0: exact table below 16, Stirling series above (error < 2e-12)
__device__ __forceinline__ double log_factorial(double k) {
    if (k < 16.0) {
        const double tab[16] = {0.0, 0.0, 0.69314718055994530942, 1.79175946922805500081, 3.17805383034794561965,
                                4.78749174278204599425, 6.57925121201010099506, 8.52516136106541430017,
                                10.60460290274525022842, 12.80182748008146961121, 15.10441257307551529523,
                                17.50230784587388583929, 19.98721449566188614952, 22.55216385312342288557,
                                25.19122118273868150009, 27.89927138384089156609};
        return tab[(int)k];
    }
    const double x = k + 1.0, r = 1.0 / x, r2 = r * r;
    return (x - 0.5) * log(x) - x + 0.91893853320467274178 + r * (1.0 / 12.0 - r2 * (1.0 / 360.0 - r2 * (1.0 / 1260.0)));
}

// Poisson with a per-call mean (Newsvendor: mu differs per env).
//
// Main path, mu < 4 * ntab: Poisson(mu) = Poisson(mu0) + Poisson(r) with mu0 = 4*floor(mu/4) and r = mu - mu0 in [0, 4)
// (a sum of independent Poisson variables is Poisson).  The first term is ONE draw from the Walker alias table of the
// tabulated mean mu0 (tables built in float64 on the host for every multiple of 4, L2-resident: one 8-byte gather per
// draw); the second is CDF inversion by chop-down search from one 32-bit uniform (p_k = p_{k-1} * r / k with a
// reciprocal table; r < 4 keeps the search at ~2 iterations on average, <= ~10 for the slowest lane of a warp).  One
// Philox block serves two periods.  Compared with a rejection sampler there is no float64 logarithm, no division and
// no retry loop -- a warp no longer runs the slow path of its unluckiest lane.
// Fallback, mu >= 4 * ntab (only reachable with fixed parameters beyond mu_max, or mu_max > 4096): Hoermann's PTRS
// transformed rejection in float64 (the algorithm numpy uses), kept out of line.
struct PoisTabDev {
    const uint2* tab;  // [ntab][1 << log2k] {threshold (u32), value if accepted | alias value << 16}
    int log2k, ntab;
};
#define ORGYM_PT_G 4
#define ORGYM_PT_MAX_MEAN 4096.0
// tables of `device` covering means up to at least mu_hi (built on first use, cached for the life of the process)
int orgym_poisson_tables(int device, double mu_hi, PoisTabDev* out);

struct PoisSplit {
    int i0;     // table index: mu0 = 4 * i0 (-1: mean beyond the tables -> PTRS)
    double r;   // mu - mu0
    double p0;  // exp(-r)
};
__device__ __forceinline__ PoisSplit poisson_split(const PoisTabDev& T, double mu) {
    PoisSplit s;
    s.i0 = 0; s.r = 0.0; s.p0 = 1.0;
    if (!(mu > 0.0)) return s;
    if (!(mu < (double)(ORGYM_PT_G * T.ntab))) { s.i0 = -1; return s; }
    s.i0 = (int)(mu * (1.0 / ORGYM_PT_G));
    s.r = mu - (double)(ORGYM_PT_G * s.i0);  // exact: both operands are within a factor 2 or r == mu
    s.p0 = exp(-s.r);
    return s;
}
// smallest k with u <= cdf_k of Poisson(r), r < 4, by sequential search: cdf_0 = p0 = exp(-r), p_k = p_{k-1} * r / k
__device__ __forceinline__ int poisson_small_invert(double u, double r, double p0, const double* __restrict__ rcp) {
    double p = p0, c = p0;
    int k = 0;
    while (u > c && k < 64) {  // P(k > 64 | r < 4) < 1e-50
        k += 1;
        p *= r * rcp[k];
        c += p;
    }
    return k;
}
// wa: word for the alias draw, wi: word for the inversion
__device__ __forceinline__ int poisson_tab_draw(const PoisTabDev& T, const double* __restrict__ rcp, const PoisSplit& s,
                                                uint32_t wa, uint32_t wi) {
    int x = 0;
    if (s.i0 > 0) {
        const uint2 e = T.tab[((size_t)s.i0 << T.log2k) + (wa >> (32 - T.log2k))];
        x = (wa << T.log2k) < e.x ? (int)(e.y & 0xFFFFu) : (int)(e.y >> 16);
    }
    return x + poisson_small_invert(((double)wi + 0.5) * (1.0 / 4294967296.0), s.r, s.p0, rcp);
}
// The rollout kernels draw T times from the same Poisson(r): they tabulate cdf_0..cdf_15 once per episode (same
// recurrence, same order -> the same doubles as poisson_small_invert sees) and replace the sequential search by a
// 4-step branch-free binary search; beyond the table (P < 1e-6 for r < 4) they fall back to the sequential search.
#define ORGYM_PT_CDF 16
__device__ __forceinline__ void poisson_small_table(double r, double p0, const double* __restrict__ rcp, double* cdf,
                                                    int stride) {
    double p = p0, c = p0;
    cdf[0] = c;
#pragma unroll
    for (int k = 1; k < ORGYM_PT_CDF; k++) {
        p *= r * rcp[k];
        c += p;
        cdf[k * stride] = c;
    }
}
__device__ __forceinline__ int poisson_small_lookup(double u, const double* cdf, int stride, double r, double p0,
                                                    const double* __restrict__ rcp) {
    int k = 0;
    k += (u > cdf[(k + 7) * stride]) ? 8 : 0;
    k += (u > cdf[(k + 3) * stride]) ? 4 : 0;
    k += (u > cdf[(k + 1) * stride]) ? 2 : 0;
    k += (u > cdf[k * stride]) ? 1 : 0;
    if (k == ORGYM_PT_CDF - 1 && u > cdf[(ORGYM_PT_CDF - 1) * stride]) return poisson_small_invert(u, r, p0, rcp);
    return k;
}
// alias part of poisson_tab_draw (the table gather), for callers that invert through poisson_small_lookup
__device__ __forceinline__ int poisson_tab_alias(const PoisTabDev& T, const PoisSplit& s, uint32_t wa) {
    if (s.i0 <= 0) return 0;
    const uint2 e = T.tab[((size_t)s.i0 << T.log2k) + (wa >> (32 - T.log2k))];
    return (wa << T.log2k) < e.x ? (int)(e.y & 0xFFFFu) : (int)(e.y >> 16);
}

struct PoissonMu {
    double mu, b, a, vr, loglam, inv_alpha;
};
// full = false: only what the squeeze (fast acceptance) needs; the rest is computed on demand in the slow path
template <bool FULL>
__device__ __forceinline__ PoissonMu ptrs_setup(double mu) {
    PoissonMu c;
    c.mu = mu;
    c.b = 0.931 + 2.53 * sqrt(mu);
    c.a = -0.059 + 0.02483 * c.b;
    c.vr = 0.9277 - 3.6224 / (c.b - 2.0);
    c.loglam = c.inv_alpha = 0.0;
    if (FULL) {
        c.loglam = log(mu);
        c.inv_alpha = 1.1239 + 1.1328 / (c.b - 3.4);
    }
    return c;
}
// mu >= 10.  Keyed by (key, episode, t); rejection attempts advance counter word 3.
template <bool FULL>
__device__ __forceinline__ int64_t ptrs_draw(const PoissonMu& c, uint64_t key, uint32_t episode, int t) {
    const double mu = c.mu;
    // Each rejection attempt consumes two 32-bit uniforms (U and V are only compared / passed through smooth
    // functions, so 2^-32 resolution is far below any statistical visibility); one Philox block feeds two attempts.
    uint4 w = make_uint4(0, 0, 0, 0);
    for (uint32_t attempt = 0;; attempt++) {
        if ((attempt & 1u) == 0) w = philox_block(key, (uint32_t)t, episode, STREAM_POISSON_MU, attempt >> 1);
        const uint32_t wu = (attempt & 1u) ? w.z : w.x, wv = (attempt & 1u) ? w.w : w.y;
        const double U = ((double)wu + 0.5) * (1.0 / 4294967296.0) - 0.5, V = ((double)wv + 0.5) * (1.0 / 4294967296.0);
        const double us = 0.5 - fabs(U);
        double kf = floor((2.0 * c.a * __drcp_rn(us) + c.b) * U + mu + 0.43);
        if (us >= 0.07 && V <= c.vr) return (int64_t)kf;  // squeeze: ~86 % of the draws end here
        if (kf < 0.0 || (us < 0.013 && V > us)) continue;
        // acceptance test  log(V*inv_alpha/(a/us^2+b)) <= -mu + k*log(mu) - log(k!).  It is first evaluated with
        // single-precision logarithms (a few instructions each); only when the two sides are closer than a rigorous
        // bound on that evaluation's error is the float64 expression computed, so the decision is always the float64 one.
        const double inv_alpha = FULL ? c.inv_alpha : 1.1239 + 1.1328 * __drcp_rn(c.b - 3.4);
        const double arg = V * inv_alpha * __drcp_rn(c.a * __drcp_rn(us * us) + c.b);
        {
            const double x1 = kf + 1.0;
            const double lf = kf < 16.0 ? log_factorial(kf)
                                        : (x1 - 0.5) * (double)__logf((float)x1) - x1 + 0.91893853320467274178 +
                                              (1.0 / 12.0) * (double)__frcp_rn((float)x1);
            const double lhs = (double)__logf((float)arg);
            const double rhs = -mu + kf * (double)__logf((float)mu) - lf;
            const double tol = 3.0e-5 * (kf + 16.0);  // >= 10x the worst-case error of the three __logf calls
            if (lhs < rhs - tol) return (int64_t)kf;
            if (lhs > rhs + tol) continue;
        }
        const double loglam = FULL ? c.loglam : log(mu);
        if (log(arg) <= (-mu + kf * loglam - log_factorial(kf))) return (int64_t)kf;
        if (attempt > 1000u) return (int64_t)kf;  // unreachable in practice; bounds the loop
    }
}
static __device__ __noinline__ int64_t poisson_ptrs(double mu, uint64_t key, uint32_t episode, int t) {
    PoissonMu c = ptrs_setup<false>(mu);
    return ptrs_draw<false>(c, key, episode, t);
}
// demand of (key, episode, period t) for mean mu -- the same value through every API (step, rollout, sampler)
__device__ __forceinline__ uint2 poisson_words(uint64_t key, uint32_t episode, int t) {
    const uint4 w = philox_block(key, (uint32_t)t >> 1, episode, STREAM_POISSON_TAB, 0);
    return (t & 1) ? make_uint2(w.z, w.w) : make_uint2(w.x, w.y);
}
__device__ __forceinline__ int64_t poisson_mu(const PoisTabDev& T, const double* __restrict__ rcp, double mu, uint64_t key,
                                              uint32_t episode, int t) {
    const PoisSplit s = poisson_split(T, mu);
    if (s.i0 < 0) return poisson_ptrs(mu, key, episode, t);
    const uint2 w = poisson_words(key, episode, t);
    return poisson_tab_draw(T, rcp, s, w.x, w.y);
}

