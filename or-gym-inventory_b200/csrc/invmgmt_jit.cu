// invmgmt_jit.cu -- run-time specialisation of the serial multi-echelon env's fused rollout.
//
// The ahead-of-time kernel (invmgmt.cu: inv_rollout_kernel) keeps the lead-time rings in shared memory and indexes
// them with run-time positions; ncu attributes ~30 % of its 216 warp-instructions per env-step to that bookkeeping
// (ring address arithmetic, LDS/STS, position wrap-around, the policy switch, the period loop, Philox word selection).
// For a given configuration all of it is static: lead times, capacities, prices and the horizon are known at create
// time, and the ring position of period t is t mod L_i.  This generator therefore emits the T periods as straight-line
// CUDA -- every ring slot a named register, every index a literal, prices / discount factors as hexadecimal float
// literals -- and compiles it for sm_100a with NVRTC (jit.cu).  The arithmetic (integer dynamics, float64 reward terms
// in the reference's order, numpy's summation order, Philox streams) is the ahead-of-time kernel's statement by
// statement; tests/test_invmgmt_gpu.py checks the two kernels bit for bit against each other.
//
// Reference: inventory_management.py:224-352 (step), benchmark_InvManagementBacklogEnv.py:142-198 (BaseStockAgent).
#include "invmgmt_jit.cuh"

#include <cmath>
#include <cstdarg>

#include "jit.cuh"

namespace {
struct Src {
    std::string s;
    void operator()(const char* fmt, ...) {
        char buf[1024];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        s += buf;
        s += '\n';
    }
};
std::string lit(double v) {
    char buf[64];
    if (v == 0.0) return "0.0";
    snprintf(buf, sizeof(buf), "%a", v);
    return buf;
}
int env_int(const char* name, int dflt, int lo, int hi) {
    const char* v = getenv(name);
    if (!v) return dflt;
    const int x = atoi(v);
    return x < lo || x > hi ? dflt : x;
}

// policy: 0 = base-stock with integer levels, 1 = uniform random orders on {0..c_i}
void emit_kernel(Src& o, const InvJitSpec& S, const char* name, int policy, int min_blocks) {
    const int n = S.n, T = S.T;
    const bool bs = policy == 0, bl = S.backlog != 0;
    const bool exact = inv_jit_profit_is_exact(S) && env_int("ORGYM_INV_JIT_FMA", 1, 0, 1) != 0;
    const char* W4[4] = {"w.x", "w.y", "w.z", "w.w"};
    const char* A4[4] = {"a4.x", "a4.y", "a4.z", "a4.w"};
    o("extern \"C\" __global__ void __launch_bounds__(NTHR, %d) %s(const InvJitArgs A) {", min_blocks, name);
    o("  __shared__ uint2 tab[%d];", 1 << S.log2k);
    o("  const int tid = threadIdx.x;");
    o("  const long long e = (long long)blockIdx.x * NTHR + tid;");
    o("  const bool valid = e < A.N;");
    o("  for (int i = tid; i < %d; i += NTHR) tab[i] = A.table[i];", 1 << S.log2k);
    o("  __syncthreads();");
    o("  const unsigned long long key = A.seed + (unsigned long long)(A.env_offset + e);");
    o("  const unsigned int ep = A.episode;");
    for (int i = 0; i < n; i++) o("  int I_%d = %lld;", i, S.I0[i]);
    if (bl)
        for (int j = 0; j <= n; j++) o("  int B_%d = 0;", j);
    for (int i = 0; i < n; i++) {
        if (bs) o("  int ps_%d = 0; const int tg_%d = A.target[%d];", i, i, i);
        for (int s = 0; s < S.L[i]; s++) {
            o("  int rr%d_%d = 0;", i, s);
            if (bs) o("  int ar%d_%d = 0;", i, s);
        }
    }
    o("  double ret = 0.0; int s_sales = 0, s_dem = 0, s_stock = 0, s_inv = 0;");
    o("  uint4 w = make_uint4(0u, 0u, 0u, 0u);");
    for (int t = 0; t < T; t++) {
        o("  {  // ---- period %d", t);
        // policy
        if (bs) {
            for (int i = 0; i < n; i++) {
                o("    int q_%d = tg_%d - (I_%d + ps_%d); q_%d = q_%d > 0 ? q_%d : 0; q_%d = q_%d < %lld ? q_%d : %lld;", i, i, i,
                  i, i, i, i, i, i, S.c[i], i, S.c[i]);
            }
        } else {
            o("    uint4 a4;");
            for (int i = 0; i < n; i++) {
                if ((i & 3) == 0) o("    a4 = philox_block(key, %uu, ep, STREAM_ACTION, %uu);", (unsigned)t, (unsigned)(i >> 2));
                o("    const int q_%d = (int)mulhi32(%s, %lluu);", i, A4[i & 3], (unsigned long long)(S.c[i] + 1));
            }
        }
        // demand
        if ((t & 3) == 0) o("    w = philox_block(key, %uu, ep, STREAM_DEMAND, 0u);", (unsigned)(t >> 2));
        o("    const int dl = alias_draw(tab, %d, %d, %s);", S.log2k, S.base, W4[t & 3]);
        // dynamics (:253-312)
        for (int i = 0; i < n; i++) {
            if (bl)
                o("    const int cur_%d = q_%d + B_%d;", i, i, i + 1);
            else
                o("    const int cur_%d = q_%d;", i, i);
            o("    int r_%d = cur_%d < %lld ? cur_%d : %lld;", i, i, S.c[i], i, S.c[i]);
            if (i + 1 < n) o("    r_%d = I_%d < r_%d ? I_%d : r_%d;", i, i + 1, i, i + 1, i);
        }
        for (int i = 0; i < n; i++) {
            if (S.L[i] == 0)
                o("    int Ic_%d = I_%d + r_%d;", i, i, i);
            else
                o("    int Ic_%d = I_%d + rr%d_%d;", i, i, i, t % S.L[i]);
        }
        o("    const int d = dl > 0 ? dl : 0;");
        if (bl)
            o("    const int fill = d + B_0;");
        else
            o("    const int fill = d;");
        o("    const int s0 = Ic_0 < fill ? Ic_0 : fill;");
        o("    Ic_0 -= s0;");
        for (int i = 1; i < n; i++) o("    Ic_%d -= r_%d;", i, i);
        o("    const int U_0 = fill - s0;");
        for (int i = 0; i < n; i++) o("    const int U_%d = cur_%d - r_%d;", i + 1, i, i);
        // profit (:315-321): elementwise float64, np.sum over m < 8 stages = sequential sum starting from 0.0
        if (exact) {
            // every operation is exact (inv_jit_profit_is_exact): one fused chain, (up - uc) folded into one coefficient
            o("    double pr = 0.0;");
            for (int j = n; j >= 0; j--) {
                o("    pr = fma(%s, (double)U_%d, pr);", lit(-S.kc[j]).c_str(), j);
                if (j < n) o("    { const int inv = Ic_%d > 0 ? Ic_%d : 0; pr = fma(%s, (double)inv, pr); }", j, j, lit(-S.hc[j]).c_str());
                if (j == 0)
                    o("    pr = fma(%s, (double)s0, pr);", lit(S.up[j] - S.uc[j]).c_str());
                else
                    o("    pr = fma(%s, (double)r_%d, pr);", lit(S.up[j] - S.uc[j]).c_str(), j - 1);
            }
        } else {
            for (int j = 0; j <= n; j++) {
                if (j == 0)
                    o("    double tm_%d; { const double s = (double)s0;", j);
                else
                    o("    double tm_%d; { const double s = (double)r_%d;", j, j - 1);
                o("      const double rev = %s * s, pc = %s * s, pen = %s * (double)U_%d;", lit(S.up[j]).c_str(),
                  lit(S.uc[j]).c_str(), lit(S.kc[j]).c_str(), j);
                if (j == n)
                    o("      tm_%d = (rev - pc) - pen; }", j);
                else
                    o("      const int inv = Ic_%d > 0 ? Ic_%d : 0; const double hold = %s * (double)inv; tm_%d = ((rev - pc) - hold) - pen; }",
                      j, j, lit(S.hc[j]).c_str(), j);
            }
            o("    double pr = 0.0;");
            for (int j = 0; j <= n; j++) o("    pr = pr + tm_%d;", j);
        }
        o("    ret += %s * pr;", lit(S.disc[(size_t)t]).c_str());
        // statistics and state
        o("    s_sales += s0; s_dem += d; s_stock += U_0;");
        for (int i = 0; i < n; i++) o("    I_%d = Ic_%d; s_inv += I_%d > 0 ? I_%d : 0;", i, i, i, i);
        if (bl)
            for (int j = 0; j <= n; j++) o("    B_%d = U_%d;", j, j);
        // rings: slot t mod L_i held R[t - L_i] / the request of period t - L_i
        for (int i = 0; i < n; i++) {
            if (S.L[i] == 0) continue;
            const int s = t % S.L[i];
            if (bs) o("    ps_%d = ps_%d + q_%d - ar%d_%d; ar%d_%d = q_%d;", i, i, i, i, s, i, s, i);
            o("    rr%d_%d = r_%d;", i, s, i);
        }
        o("  }");
    }
    // outputs: identical to the ahead-of-time kernel
    o("  if (valid) {");
    o("    if (A.ep_return) A.ep_return[e] = ret;");
    o("    if (A.stats) *reinterpret_cast<longlong4*>(A.stats + e * 4) = make_longlong4((long long)s_sales, (long long)s_dem, (long long)s_stock, (long long)s_inv);");
    o("    if (A.stats32) *reinterpret_cast<int4*>(A.stats32 + e * 4) = make_int4(s_sales, s_dem, s_stock, s_inv);");
    o("  }");
    o("  if (A.partials) {");
    o("    double v[7];");
    o("    v[0] = valid ? 1.0 : 0.0; v[1] = valid ? ret : 0.0; v[2] = valid ? ret * ret : 0.0;");
    o("    v[3] = valid ? (double)s_sales : 0.0; v[4] = valid ? (double)s_dem : 0.0;");
    o("    v[5] = valid ? (double)s_stock : 0.0; v[6] = valid ? (double)s_inv : 0.0;");
    o("    __shared__ double red[NTHR / 32][7];");
    o("    _Pragma(\"unroll\") for (int q = 0; q < 7; q++) { double x = v[q];");
    o("      _Pragma(\"unroll\") for (int s = 16; s > 0; s >>= 1) x += __shfl_xor_sync(0xffffffffu, x, s);");
    o("      if ((tid & 31) == 0) red[tid >> 5][q] = x; }");
    o("    __syncthreads();");
    o("    if (tid < 7) { double x = 0.0; for (int wv = 0; wv < NTHR / 32; wv++) x += red[wv][tid]; A.partials[(long long)blockIdx.x * 8 + tid] = x; }");
    o("  }");
    o("}");
}
}  // namespace

// The profit of a period is a sum of products coefficient x integer.  Every coefficient is a float32 value, i.e. an
// integer multiple of some power of two; let Q be the smallest such quantum over all coefficients.  Then every product,
// partial sum and difference that can occur is an integer multiple of Q, and float64 represents all multiples of Q below
// 2^53 * Q exactly.  If the largest possible magnitude  xbound * sum_j (|up_j| + |uc_j| + |hc_j| + |kc_j|)  stays below
// that, no operation ever rounds: the reference's evaluation order, any re-association and fused multiply-adds all give
// the same float64 value, and the generator may use the cheapest form (3 DFMA per stage instead of 4 DMUL + 3 DADD).
bool inv_jit_profit_is_exact(const InvJitSpec& S) {
    if (S.xbound <= 0) return false;
    int qexp = 1 << 20;  // exponent of the finest quantum
    double mag = 0.0;
    for (int j = 0; j <= S.n; j++) {
        const double cs[4] = {S.up[j], S.uc[j], S.hc[j], S.kc[j]};
        for (double c : cs) {
            if (!(c >= 0.0) || !std::isfinite(c)) return false;
            mag += c;
            if (c == 0.0) continue;
            if ((double)(float)c != c) return false;  // not a float32 value: no 24-bit significand guarantee
            int e;
            double m = std::frexp(c, &e);             // c = m * 2^e, 0.5 <= m < 1
            long long mi = (long long)std::ldexp(m, 53);  // 53-bit integer significand
            int tz = 0;
            while ((mi & 1) == 0) { mi >>= 1; tz++; }
            qexp = std::min(qexp, e - 53 + tz);       // c = odd * 2^(e - 53 + tz)
        }
    }
    if (qexp == (1 << 20)) return true;               // all coefficients zero
    const double limit = std::ldexp(1.0, 52 + qexp);  // half of 2^53 * Q: one bit of head-room
    return mag * (double)S.xbound < limit;
}

bool inv_jit_eligible(const InvJitSpec& S) {
    if (S.n < 1 || S.n > 6) return false;  // m = n + 1 < 8: the reward sum is numpy's sequential branch
    if (S.T < 1 || S.T > 64 || (int)S.disc.size() != S.T) return false;
    if (S.log2k < 0 || S.log2k > 10) return false;
    int sumL = 0;
    for (int i = 0; i < S.n; i++) {
        if (S.c[i] < 0 || S.c[i] > 0x3fffffffLL || S.I0[i] < 0 || S.I0[i] > 0x3fffffffLL) return false;
        sumL += S.L[i];
    }
    return sumL <= 40;  // two register-resident rings of sum(L) entries each
}

std::string inv_jit_source(const InvJitSpec& S) {
    Src o;
    o.s += orgym_jit_device_rng_src();
    o.s += orgym_jit_inv_args_src();
    o("#define NTHR %d", INV_JIT_THREADS);
    const int mb = env_int("ORGYM_INV_JIT_MINBLOCKS", 4, 1, 8);
    emit_kernel(o, S, "inv_jit_rollout_bs", 0, mb);
    emit_kernel(o, S, "inv_jit_rollout_rnd", 1, mb);
    return o.s;
}
