// report.cu -- the reference's per-agent summary row, computed on the device from the outputs of a fused rollout.
//
// `evaluate_agent` (benchmark_InvManagementBacklogEnv.py:381-441, benchmark_NetInvMgmtBacklogEnv.py:223-303) records per
// episode TotalReward, AvgServiceLevel, TotalStockoutQty, AvgEndingInv; `process_and_report_results` (:493-504 / :320-330)
// aggregates them with pandas: mean / median / std (ddof = 1) / min / max of TotalReward, mean of the other three.  For
// 2^24 episodes per rollout that row is the product a user wants back -- not 134 MB of per-episode returns over PCIe.
//
//   pass 0 (all episodes)   sum, min, max of the returns; sums of the service level sales/demand (1.0 without demand,
//                           :425), unfulfilled quantity, on-hand inventory.  The top 11 bits (sign, exponent) of the
//                           order-preserving 64-bit image of the returns are normally the same for min and max, which
//                           fixes the first digit without a histogram (pass 0b builds one otherwise)
//   pass 1 (all episodes)   sum (x - mean)^2 (two-pass variance); histogram of the next 11 bits among the returns that
//                           share the 11-bit prefix of the wanted order statistic
//   pass 2 (all episodes)   the returns that share the 22-bit prefix are compacted into a candidate list (a few
//                           thousand of 2^24 for a continuous distribution) and histogrammed on the next 11 bits
//   passes 3-5 (candidates) remaining digits (11 + 10 + 10 bits): the exact order statistic -- radix select, no sort
// pandas' median of an even count is the midpoint of the two central order statistics: both are selected side by side
// (their prefixes usually coincide; when they do not, each keeps its own histogram and candidate list).
// Floating-point sums are per-CTA partials added in a fixed order by one CTA -> bitwise reproducible; histograms are
// integer atomics.  Everything is enqueued on the caller's stream without a host round trip.
#include "common.cuh"

namespace {
constexpr int RCTAS = 592;        // 4 CTAs per SM on a 148-SM B200: one wave
constexpr int RTHR = 256;
constexpr int NBIN = 2048;
constexpr int NPASS = 6;
constexpr int UNR = 4;          // independent loads per thread and trip in the streaming passes
__constant__ int c_shift[NPASS] = {53, 42, 31, 20, 10, 0};
__constant__ int c_bits[NPASS] = {11, 11, 11, 11, 10, 10};

struct RState {                   // device-resident control block
    unsigned long long n;
    unsigned long long k[2];      // remaining rank of the two wanted order statistics inside their prefix
    unsigned long long prefix[2]; // bits fixed so far (right-aligned)
    unsigned long long list_n[2]; // candidates appended (may exceed the capacity: then the lists are not used)
    int list_of[2];               // which list holds the candidates of selection s
    int lists_ok;                 // passes 3-5 read the candidate lists (else all episodes again)
    int need_hist0;               // the returns span several top-11-bit bins: pass 0b builds that histogram
    int same;                     // both statistics share prefix, histogram and list so far
    double mean;
    double sums[8];               // pass 0: sum x, sum sl, sum unf, sum inv, -, -, -, -; [6] = sum (x - mean)^2
    unsigned long long kmin, kmax;
};

__device__ __forceinline__ unsigned long long okey(double x) {  // order-preserving map double -> uint64
    const long long b = __double_as_longlong(x);
    return (unsigned long long)(b ^ ((b >> 63) | (long long)0x8000000000000000LL));
}
__device__ __forceinline__ double okey_inv(unsigned long long u) {
    const long long b = (u & 0x8000000000000000ULL) ? (long long)(u ^ 0x8000000000000000ULL) : (long long)~u;
    return __longlong_as_double(b);
}

// stats element (episode e, column c) as double
template <int KIND>
__device__ __forceinline__ void load_stats(const void* st, long long e, double (&v)[4]) {
    if (KIND == 0) {  // int64 [N,4]
        const longlong4 a = *reinterpret_cast<const longlong4*>((const long long*)st + e * 4);
        v[0] = (double)a.x; v[1] = (double)a.y; v[2] = (double)a.z; v[3] = (double)a.w;
    } else if (KIND == 1) {  // int32 [N,4]
        const int4 a = *reinterpret_cast<const int4*>((const int*)st + e * 4);
        v[0] = (double)a.x; v[1] = (double)a.y; v[2] = (double)a.z; v[3] = (double)a.w;
    } else {  // float64 [N,4]
        const double4 a = *reinterpret_cast<const double4*>((const double*)st + e * 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    }
}

// block-wide deterministic sum of NV doubles per thread -> partial row [blockIdx][NV] (thread 0..NV-1 write)
template <int NV>
__device__ __forceinline__ void block_sums(double (&v)[NV], double* row) {
    __shared__ double red[RTHR / 32][NV];
    const int tid = threadIdx.x;
#pragma unroll
    for (int q = 0; q < NV; q++) {
        double x = v[q];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) x += __shfl_xor_sync(0xffffffffu, x, s);
        if ((tid & 31) == 0) red[tid >> 5][q] = x;
    }
    __syncthreads();
    if (tid < NV) {
        double x = 0.0;
        for (int w = 0; w < RTHR / 32; w++) x += red[w][tid];
        row[tid] = x;
    }
}

template <int KIND>
__global__ void __launch_bounds__(RTHR) report_pass0(const double* __restrict__ ret, const void* __restrict__ stats, long long n,
                                                     double* __restrict__ partials, unsigned long long* __restrict__ kmm) {
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    unsigned long long kmin = ~0ULL, kmax = 0ULL;
    // UNR independent loads per thread and trip (the passes are pure streaming: memory-level parallelism is what counts)
    for (long long e0 = (long long)blockIdx.x * (RTHR * UNR) + threadIdx.x; e0 < n; e0 += (long long)gridDim.x * (RTHR * UNR)) {
        double x[UNR], s[UNR][4];
#pragma unroll
        for (int k = 0; k < UNR; k++) {
            const long long e = e0 + (long long)k * RTHR;
            x[k] = e < n ? ret[e] : 0.0;
            s[k][0] = s[k][1] = s[k][2] = s[k][3] = 0.0;
            if (stats && e < n) load_stats<KIND>(stats, e, s[k]);
        }
#pragma unroll
        for (int k = 0; k < UNR; k++) {
            if (e0 + (long long)k * RTHR >= n) continue;
            const unsigned long long u = okey(x[k]);
            v[0] += x[k];
            kmin = u < kmin ? u : kmin;
            kmax = u > kmax ? u : kmax;
            if (stats) {
                v[1] += s[k][1] > 1e-6 ? s[k][0] / (s[k][1] > 1e-6 ? s[k][1] : 1e-6) : 1.0;  // :425 per-episode service level
                v[2] += s[k][2];
                v[3] += s[k][3];
            }
        }
    }
    block_sums<4>(v, partials + (size_t)blockIdx.x * 8);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const unsigned long long a = __shfl_xor_sync(0xffffffffu, kmin, s), b = __shfl_xor_sync(0xffffffffu, kmax, s);
        kmin = a < kmin ? a : kmin;
        kmax = b > kmax ? b : kmax;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&kmm[0], kmin);
        atomicMax(&kmm[1], kmax);
    }
}

// pass 0b: histogram of the top 11 bits, only when min and max differ there (mixed signs, many binades)
__global__ void __launch_bounds__(RTHR) report_pass0b(const double* __restrict__ ret, long long n, const RState* __restrict__ S,
                                                      unsigned long long* __restrict__ hist) {
    if (!S->need_hist0) return;
    __shared__ unsigned int sh[NBIN];
    for (int i = threadIdx.x; i < NBIN; i += RTHR) sh[i] = 0u;
    __syncthreads();
    for (long long e = (long long)blockIdx.x * RTHR + threadIdx.x; e < n; e += (long long)gridDim.x * RTHR)
        atomicAdd(&sh[okey(ret[e]) >> 53], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < NBIN; i += RTHR)
        if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

// passes 1..5.  FULL: keys come from the returns themselves; otherwise from the candidate list of selection s.
// APPEND (pass 2): matching keys are also compacted into the candidate list.  VAR (pass 1): sum (x - mean)^2.
// Passes 3-5 are enqueued in both forms; the one that is not its turn (S->lists_ok) returns at once.
template <bool FULL, bool APPEND, bool VAR>
__global__ void __launch_bounds__(RTHR) report_pass(int pass, const double* __restrict__ ret, long long n, RState* __restrict__ S,
                                                    unsigned long long* __restrict__ hist, unsigned long long* __restrict__ list,
                                                    long long cap, double* __restrict__ partials) {
    if (pass >= 3 && (FULL ? S->lists_ok != 0 : S->lists_ok == 0)) return;
    __shared__ unsigned int sh[2][NBIN];
    constexpr unsigned int SBUF = APPEND ? 1024u : 1u;  // candidates buffered per CTA and list before they are published
    __shared__ unsigned long long sbuf[2][SBUF];
    __shared__ unsigned int scount[2];
    __shared__ unsigned long long sbase[2];
    const int same = S->same;
    const int nsel = same ? 1 : 2;
    for (int i = threadIdx.x; i < 2 * NBIN; i += RTHR) (&sh[0][0])[i] = 0u;
    if (threadIdx.x < 2) scount[threadIdx.x] = 0u;
    __syncthreads();
    const int shift = c_shift[pass], pshift = c_shift[pass - 1];
    const unsigned int mask = (1u << c_bits[pass]) - 1u;
    const unsigned long long p0 = S->prefix[0], p1 = S->prefix[1];
    double var = 0.0;
    const double mean = S->mean;
    if (FULL) {
        for (long long e0 = (long long)blockIdx.x * (RTHR * UNR) + threadIdx.x; e0 < n; e0 += (long long)gridDim.x * (RTHR * UNR)) {
          double xs[UNR];
#pragma unroll
          for (int k = 0; k < UNR; k++) xs[k] = e0 + (long long)k * RTHR < n ? ret[e0 + (long long)k * RTHR] : 0.0;
#pragma unroll
          for (int k = 0; k < UNR; k++) {
            if (e0 + (long long)k * RTHR >= n) continue;
            const double x = xs[k];
            if (VAR) {
                const double d = x - mean;
                var += d * d;
            }
            const unsigned long long u = okey(x), pre = u >> pshift;
            const bool m0 = pre == p0, m1 = !same && pre == p1;
            if (m0) atomicAdd(&sh[0][(unsigned int)(u >> shift) & mask], 1u);
            if (m1) atomicAdd(&sh[1][(unsigned int)(u >> shift) & mask], 1u);
            if (APPEND) {
                // Candidates are collected per CTA in shared memory and published with ONE global atomic per CTA and list
                // at the end (the first version appended with one global atomic per warp: ~110 000 atomics on one address,
                // 65 of this pass's 100 microseconds).  A full CTA buffer falls back to direct, warp-aggregated appends.
                bool d0 = false, d1 = false;
                if (m0) {
                    const unsigned int at = atomicAdd(&scount[0], 1u);
                    if (at < SBUF) sbuf[0][at] = u; else d0 = true;
                }
                if (m1) {
                    const unsigned int at = atomicAdd(&scount[1], 1u);
                    if (at < SBUF) sbuf[1][at] = u; else d1 = true;
                }
                const unsigned act = __activemask();
                const unsigned b0 = __ballot_sync(act, d0), b1 = __ballot_sync(act, d1);
                const int lane = threadIdx.x & 31;
                if (b0) {
                    unsigned long long base = 0ULL;
                    if (lane == __ffs(b0) - 1) base = atomicAdd(&S->list_n[0], (unsigned long long)__popc(b0));
                    base = __shfl_sync(act, base, __ffs(b0) - 1);
                    const unsigned long long at = base + __popc(b0 & ((1u << lane) - 1u));
                    if (d0 && (long long)at < cap) list[at] = u;
                }
                if (b1) {
                    unsigned long long base = 0ULL;
                    if (lane == __ffs(b1) - 1) base = atomicAdd(&S->list_n[1], (unsigned long long)__popc(b1));
                    base = __shfl_sync(act, base, __ffs(b1) - 1);
                    const unsigned long long at = base + __popc(b1 & ((1u << lane) - 1u));
                    if (d1 && (long long)at < cap) list[cap + at] = u;
                }
            }
          }
        }
        if (VAR) {
            double v[1] = {var};
            block_sums<1>(v, partials + (size_t)blockIdx.x * 8);
        }
        if (APPEND) {  // publish the CTA's candidates: one global atomic per list
            __syncthreads();
            for (int s = 0; s < 2; s++) {
                const unsigned int cnt = scount[s] < SBUF ? scount[s] : SBUF;
                if (cnt == 0) continue;  // uniform: scount is shared
                if (threadIdx.x == 0) sbase[s] = atomicAdd(&S->list_n[s], (unsigned long long)cnt);
                __syncthreads();
                for (unsigned int i = threadIdx.x; i < cnt; i += RTHR) {
                    const unsigned long long at = sbase[s] + i;
                    if ((long long)at < cap) list[(size_t)s * cap + at] = sbuf[s][i];
                }
            }
        }
    } else {
        for (int s = 0; s < nsel; s++) {
            const int li = S->list_of[s];
            const long long m = (long long)S->list_n[li];
            const unsigned long long ps = s ? p1 : p0;
            const unsigned long long* L = list + (size_t)li * cap;
            for (long long e = (long long)blockIdx.x * RTHR + threadIdx.x; e < m; e += (long long)gridDim.x * RTHR) {
                const unsigned long long u = L[e];
                if ((u >> pshift) == ps) atomicAdd(&sh[s][(unsigned int)(u >> shift) & mask], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nsel * NBIN; i += RTHR) {
        const unsigned int c = (&sh[0][0])[i];
        if (c) atomicAdd(&hist[i], (unsigned long long)c);
    }
}

// one CTA: fixed-order reduction of the partial sums (passes 0 and 1), digit pick for both selections (block-wide prefix
// scan of the histogram: every thread owns NBIN / RTHR consecutive bins), histogram reset; after the last pass: the report
__global__ void __launch_bounds__(RTHR) report_pick(int pass, int sub, long long n, int periods, int npart, const double* __restrict__ partials,
                                                    unsigned long long* __restrict__ kmm, RState* __restrict__ S,
                                                    unsigned long long* __restrict__ hist, long long cap, double* __restrict__ report) {
    __shared__ double sums[8];
    __shared__ unsigned long long wtot[RTHR / 32];
    __shared__ int s_pick, s_same;
    __shared__ int s_digit[2];
    __shared__ unsigned long long s_below[2];
    const int tid = threadIdx.x;
    if ((pass == 0 && sub == 0) || pass == 1) {
        // partial rows -> sums in a fixed order: 64 threads per column add a strided subset of the rows, then the 64
        // partial results are added in thread order (bitwise reproducible for a given grid size)
        __shared__ double part[4][64];
        const int ncol = pass == 0 ? 4 : 1;
        const int col = tid >> 6, sub64 = tid & 63;
        if (col < ncol) {
            double x = 0.0;
            for (int r = sub64; r < npart; r += 64) x += partials[(size_t)r * 8 + col];
            part[col][sub64] = x;
        }
        __syncthreads();
        if (tid < ncol) {
            double x = 0.0;
            for (int q = 0; q < 64; q++) x += part[tid][q];
            sums[tid] = x;
        }
        __syncthreads();
    }
    if (tid == 0) {
        bool pick = true;
        if (pass == 0 && sub == 0) {
            S->n = (unsigned long long)n;
            for (int q = 0; q < 4; q++) S->sums[q] = sums[q];
            S->mean = sums[0] / (double)n;
            S->kmin = kmm[0];
            S->kmax = kmm[1];
            S->k[0] = (unsigned long long)((n - 1) / 2);  // pandas median: midpoint of the two central order statistics
            S->k[1] = (unsigned long long)(n / 2);
            S->prefix[0] = S->prefix[1] = 0ULL;
            S->list_n[0] = S->list_n[1] = 0ULL;
            S->list_of[0] = S->list_of[1] = 0;
            S->lists_ok = 0;
            S->same = 1;
            S->need_hist0 = (S->kmin >> 53) != (S->kmax >> 53);
            if (!S->need_hist0) S->prefix[0] = S->prefix[1] = S->kmin >> 53;  // every return shares the first digit
            pick = false;
        } else if (pass == 0) {
            pick = S->need_hist0 != 0;
        } else if (pass == 1)
            S->sums[6] = sums[0];
        s_pick = pick ? 1 : 0;
        s_same = S->same;
        s_digit[0] = s_digit[1] = (1 << c_bits[pass]) - 1;
        s_below[0] = s_below[1] = 0ULL;
    }
    __syncthreads();
    const int same = s_same;
    if (s_pick) {
        constexpr int BPT = NBIN / RTHR;  // bins per thread
        for (int s = 0; s < 2; s++) {
            const unsigned long long* h = hist + (size_t)((same ? 0 : s) * NBIN);
            const unsigned long long k = S->k[s];
            unsigned long long v[BPT], tot = 0ULL;
#pragma unroll
            for (int b = 0; b < BPT; b++) {
                v[b] = h[tid * BPT + b];
                tot += v[b];
            }
            unsigned long long inc = tot;  // inclusive scan of the per-thread totals: warp shuffles, then across warps
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long y = __shfl_up_sync(0xffffffffu, inc, o);
                if ((tid & 31) >= o) inc += y;
            }
            if ((tid & 31) == 31) wtot[tid >> 5] = inc;
            __syncthreads();
            unsigned long long base = 0ULL;
            for (int w = 0; w < (tid >> 5); w++) base += wtot[w];
            unsigned long long below = base + inc - tot;  // returns in the bins before this thread's first bin
            if (k >= below && k < below + tot) {          // the wanted rank falls into one of this thread's bins
#pragma unroll
                for (int b = 0; b < BPT; b++) {
                    if (k < below + v[b]) {
                        s_digit[s] = tid * BPT + b;
                        s_below[s] = below;
                        break;
                    }
                    below += v[b];
                }
            }
            __syncthreads();
        }
    }
    if (tid == 0) {
        if (s_pick) {
            for (int s = 0; s < 2; s++) {
                S->k[s] -= s_below[s];
                S->prefix[s] = (S->prefix[s] << c_bits[pass]) | (unsigned long long)s_digit[s];
            }
            if (S->prefix[0] != S->prefix[1]) S->same = 0;
        }
        if (pass == 2) {  // pass 2 compacted the candidates: shared list while the selections coincided, else one each
            S->list_of[0] = 0;
            S->list_of[1] = same ? 0 : 1;
            S->lists_ok = (long long)S->list_n[0] <= cap && (same || (long long)S->list_n[1] <= cap);
        }
        if (pass == NPASS - 1) {
            const double x0 = okey_inv(S->prefix[0]), x1 = okey_inv(S->prefix[1]);
            const double nn = (double)n;
            report[0] = nn;
            report[1] = S->mean;
            report[2] = 0.5 * (x0 + x1);
            report[3] = n > 1 ? sqrt(S->sums[6] / (nn - 1.0)) : __longlong_as_double(0x7ff8000000000000LL);  // ddof = 1
            report[4] = okey_inv(S->kmin);
            report[5] = okey_inv(S->kmax);
            report[6] = S->sums[1] / nn;                                    // AvgServiceLevel
            report[7] = S->sums[2] / nn;                                    // AvgStockoutQty
            report[8] = S->sums[3] / (double)(periods > 0 ? periods : 1) / nn;  // AvgEndInv
            report[9] = S->sums[0];
            report[10] = S->sums[6];
            report[11] = (double)S->list_n[0];
        }
    }
    __syncthreads();
    for (int i = tid; i < 2 * NBIN; i += RTHR) hist[i] = 0ULL;
}

struct Layout {
    double* partials;             // [RCTAS][8]
    unsigned long long* kmm;      // [2]
    unsigned long long* hist;     // [2][NBIN]
    RState* S;
    unsigned long long* list;     // [2][cap]
    long long cap;
};
long long list_cap(long long n) { return n / 16 > 65536 ? n / 16 : 65536; }
size_t fixed_bytes() { return (size_t)RCTAS * 8 * 8 + 16 + (size_t)2 * NBIN * 8 + ((sizeof(RState) + 255) & ~(size_t)255); }
Layout layout(void* scratch, long long n) {
    Layout L;
    char* p = (char*)scratch;
    L.partials = (double*)p;
    p += (size_t)RCTAS * 8 * 8;
    L.kmm = (unsigned long long*)p;
    p += 16;
    L.hist = (unsigned long long*)p;
    p += (size_t)2 * NBIN * 8;
    L.S = (RState*)p;
    p += (sizeof(RState) + 255) & ~(size_t)255;
    L.list = (unsigned long long*)p;
    L.cap = list_cap(n);
    return L;
}
}  // namespace

extern "C" int64_t orgym_report_scratch_bytes(int64_t num_episodes) {
    if (num_episodes < 0) return -1;
    return (int64_t)(fixed_bytes() + (size_t)2 * list_cap(num_episodes) * 8);
}

extern "C" int orgym_evaluation_report(int device, const double* ep_return_dev, const void* stats_dev, int stats_kind,
                                       int64_t num_episodes, int32_t periods, void* scratch_dev, double* report_dev,
                                       void* stream) {
    ORGYM_REQUIRE(ep_return_dev && scratch_dev && report_dev, "ep_return_dev, scratch_dev and report_dev are required");
    ORGYM_REQUIRE(num_episodes > 0, "no episodes");
    ORGYM_REQUIRE(stats_kind >= 0 && stats_kind <= 2, "stats_kind must be 0 (int64), 1 (int32) or 2 (float64)");
    ORGYM_REQUIRE(((uintptr_t)scratch_dev & 255) == 0, "scratch_dev must be 256-byte aligned");
    if (orgym_device_count() <= 0) {
        orgym_set_error("no CUDA device");
        return ORGYM_E_CUDA;
    }
    DeviceGuard g(device);
    cudaStream_t s = (cudaStream_t)stream;
    const long long n = num_episodes;
    Layout L = layout(scratch_dev, n);
    int grid = (int)((n + RTHR * UNR - 1) / (RTHR * UNR));
    if (grid > RCTAS) grid = RCTAS;
    // kmin = all ones, kmax = 0, histograms and list counters zero
    ORGYM_CUDA(cudaMemsetAsync(L.kmm, 0xff, 8, s));
    ORGYM_CUDA(cudaMemsetAsync(L.kmm + 1, 0, 8, s));
    ORGYM_CUDA(cudaMemsetAsync(L.hist, 0, (size_t)2 * NBIN * 8, s));
    // pass 0 holds UNR returns + UNR statistics rows per thread (77-80 registers -> 3 CTAs per SM): its grid is one wave of that
    const int grid0 = grid < 444 ? grid : 444;
    if (stats_kind == 0)
        report_pass0<0><<<grid0, RTHR, 0, s>>>(ep_return_dev, stats_dev, n, L.partials, L.kmm);
    else if (stats_kind == 1)
        report_pass0<1><<<grid0, RTHR, 0, s>>>(ep_return_dev, stats_dev, n, L.partials, L.kmm);
    else
        report_pass0<2><<<grid0, RTHR, 0, s>>>(ep_return_dev, stats_dev, n, L.partials, L.kmm);
    report_pick<<<1, RTHR, 0, s>>>(0, 0, n, periods, grid0, L.partials, L.kmm, L.S, L.hist, L.cap, report_dev);
    report_pass0b<<<grid, RTHR, 0, s>>>(ep_return_dev, n, L.S, L.hist);
    report_pick<<<1, RTHR, 0, s>>>(0, 1, n, periods, grid, L.partials, L.kmm, L.S, L.hist, L.cap, report_dev);
    report_pass<true, false, true><<<grid, RTHR, 0, s>>>(1, ep_return_dev, n, L.S, L.hist, L.list, L.cap, L.partials);
    report_pick<<<1, RTHR, 0, s>>>(1, 0, n, periods, grid, L.partials, L.kmm, L.S, L.hist, L.cap, report_dev);
    report_pass<true, true, false><<<grid, RTHR, 0, s>>>(2, ep_return_dev, n, L.S, L.hist, L.list, L.cap, L.partials);
    report_pick<<<1, RTHR, 0, s>>>(2, 0, n, periods, grid, L.partials, L.kmm, L.S, L.hist, L.cap, report_dev);
    for (int pass = 3; pass < NPASS; pass++) {
        // enqueued in both forms (candidate lists / all episodes); the kernels themselves look at S->lists_ok, so no host
        // round trip is needed to choose
        report_pass<true, false, false><<<grid, RTHR, 0, s>>>(pass, ep_return_dev, n, L.S, L.hist, L.list, L.cap, L.partials);
        report_pass<false, false, false><<<grid, RTHR, 0, s>>>(pass, ep_return_dev, n, L.S, L.hist, L.list, L.cap, L.partials);
        report_pick<<<1, RTHR, 0, s>>>(pass, 0, n, periods, grid, L.partials, L.kmm, L.S, L.hist, L.cap, report_dev);
    }
    ORGYM_CUDA(cudaGetLastError());
    return ORGYM_OK;
}
