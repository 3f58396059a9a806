// common.cuh -- shared host/device utilities of liborgym_b200.so (sm_100a only).
//
//  * error plumbing for the C ABI (thread-local last error, CUDA checks)
//  * Philox4x32-10 counter-based generator and the stream/counter conventions
//  * alias-table sampling for the fixed demand distributions, PTRS/inversion for per-env Poisson means
//  * 1-D bulk async copies (TMA engine: cp.async.bulk, SASS UBLKCP) used to move the row-major
//    [env][obs_dim] / [env][act_dim] tiles of the Gymnasium-facing tensors between HBM and shared memory
//  * numpy-order summation helpers
//
// The whole library is compiled with --fmad=false: the reference (numpy / CPython) never contracts a*b+c,
// and rewards are checked bit-for-bit against it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/orgym_b200.h"

// ------------------------------------------------------------------------------------------------
// host: errors
// ------------------------------------------------------------------------------------------------
void orgym_set_error(const char* fmt, ...);

#define ORGYM_CUDA(call)                                                                      \
    do {                                                                                      \
        cudaError_t _e = (call);                                                              \
        if (_e != cudaSuccess) {                                                              \
            orgym_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return ORGYM_E_CUDA;                                                              \
        }                                                                                     \
    } while (0)

#define ORGYM_REQUIRE(cond, ...)        \
    do {                                \
        if (!(cond)) {                  \
            orgym_set_error(__VA_ARGS__); \
            return ORGYM_E_INVALID;     \
        }                               \
    } while (0)

enum { FAM_INVMGMT = 1, FAM_NEWSVENDOR = 2, FAM_NETINV = 3 };

// every family handle starts with this header (orgym_errors / destroy dispatch on it)
struct HandleBase {
    uint32_t magic;  // 'ORGY'
    int family;
    int device;
    int64_t num_envs;
    uint32_t* err_dev;  // sticky device error bits
};
#define ORGYM_MAGIC 0x4F524759u

int orgym_handle_base_init(HandleBase* b, int family, int device, int64_t num_envs);
void orgym_handle_base_free(HandleBase* b);
int orgym_check_handle(const void* h, int family);

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// ------------------------------------------------------------------------------------------------
// alias tables (host builder in api.cu; AliasDev and the device-side samplers live in device_rng.cuh)
// ------------------------------------------------------------------------------------------------
#include "device_rng.cuh"

// builds the table on the host and uploads it; the cudaMalloc'd pointers are appended to *allocs
int orgym_build_alias(const orgym_dist_t* d, int user_clamp, AliasDev* out, std::vector<void*>* allocs);
// host pmf used by the builder
int orgym_dist_pmf(const orgym_dist_t* d, std::vector<double>* pmf, int64_t* base);

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
// wa: word for the alias draw, wi: word for the inversion
__device__ __forceinline__ int poisson_tab_draw(const PoisTabDev& T, const double* __restrict__ rcp, const PoisSplit& s,
                                                uint32_t wa, uint32_t wi) {
    int x = 0;
    if (s.i0 > 0) {
        const uint2 e = T.tab[((size_t)s.i0 << T.log2k) + (wa >> (32 - T.log2k))];
        x = (wa << T.log2k) < e.x ? (int)(e.y & 0xFFFFu) : (int)(e.y >> 16);
    }
    double u = ((double)wi + 0.5) * (1.0 / 4294967296.0), p = s.p0;
    int k = 0;
    while (u > p && k < 64) {  // P(k > 64 | r < 4) < 1e-50
        u -= p;
        k += 1;
        p *= s.r * rcp[k];
    }
    return x + k;
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

// ------------------------------------------------------------------------------------------------
// device: numpy summation order (np.add.reduce pairwise: sequential below 8, 8 partial sums up to 128)
// ------------------------------------------------------------------------------------------------
template <typename T, int CAP>
__device__ __forceinline__ T np_sum_reg(const T (&a)[CAP], int n) {
    if (n < 8) {
        T res = (T)0;
#pragma unroll
        for (int i = 0; i < CAP; i++)
            if (i < n && i < 7) res = res + a[i];
        return res;
    }
    T r[8];
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = a[i < CAP ? i : 0];
    int full = n - (n % 8);
#pragma unroll
    for (int i = 8; i < CAP; i++)
        if (i < full) r[i % 8] = r[i % 8] + a[i];
    T res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
#pragma unroll
    for (int i = 8; i < CAP; i++)
        if (i >= full && i < n) res = res + a[i];
    return res;
}

// ------------------------------------------------------------------------------------------------
// device: 1-D bulk async copies through the TMA engine (cp.async.bulk; SASS UBLKCP) + mbarrier
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait: a lost transaction traps (error return) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; spin++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (!done && spin > (1u << 24)) __trap();
    }
}
// global -> shared, completion signalled on an mbarrier.  dst/src 16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global (bulk group)
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy before a bulk store reads them
// per-thread asynchronous 4-byte copy global -> shared (SASS LDGSTS): no destination register, so the issuing thread
// keeps going until cp_async_wait_all()
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#define ORGYM_TILE 128  // env instances per CTA in the step kernels

// Deterministic reduction of per-CTA partial sums ([nrows][8] float64, columns 0..6 used) into out[8]:
// up to 128 CTAs reduce contiguous row ranges in a fixed order, the last CTA to finish (ticket) adds their results
// in CTA order.  `partials` must have room for ORGYM_REDUCE_EXTRA more doubles after the nrows*8 entries.
#define ORGYM_REDUCE_CTAS 128
#define ORGYM_REDUCE_EXTRA (ORGYM_REDUCE_CTAS * 8 + 8)
int orgym_launch_reduce(double* partials, int nrows, double* out_dev, cudaStream_t s);

static inline int64_t round_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
