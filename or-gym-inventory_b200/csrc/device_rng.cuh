// device_rng.cuh -- device-side random numbers shared by the ahead-of-time kernels (via common.cuh) and by the
// run-time specialised kernels (embedded as text and compiled with NVRTC, see jit.cu): Philox4x32-10, the stream /
// counter conventions, alias-table draws for the fixed demand distributions.  No host headers in here.
#pragma once
#ifdef __CUDACC_RTC__
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef short int16_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#else
#include <stdint.h>
#endif
#ifndef ORGYM_DIST_USER
#define ORGYM_DIST_USER 5
#endif

struct AliasDev {
    const uint2* table;  // [1 << log2k] {threshold (u32, accept if u < thr), alias index}
    int32_t log2k;
    int32_t base;        // sample = base + index
    int32_t kind;        // ORGYM_DIST_*; ORGYM_DIST_USER -> trace lookup
    const int64_t* user_D;
    int32_t user_D_len;
    int32_t user_clamp;  // 1: index min(t, len-1) (network env), 0: 0 beyond the end (serial env)
};

// ------------------------------------------------------------------------------------------------
// device: Philox4x32-10
// ------------------------------------------------------------------------------------------------
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

// counter word 2 = stream id
enum {
    STREAM_DEMAND = 0,    // c0 = period >> 2 (four alias samples per block), c3 = demand source (retail link) index
    STREAM_ACTION = 1,    // random-action policy: c0 = period, c3 = stage group (4 stages per block)
    STREAM_PARAMS = 2,    // newsvendor reset uniforms: c0 = 0..2
    STREAM_POISSON_MU = 3,  // per-env-mean Poisson, PTRS fallback: c0 = period, c3 = attempt pair
    STREAM_POISSON_TAB = 4  // per-env-mean Poisson, table + inversion: c0 = period >> 1 (two words per period)
};

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = mulhi32(PHILOX_M0, c.x), lo0 = PHILOX_M0 * c.x;
        uint32_t hi1 = mulhi32(PHILOX_M1, c.z), lo1 = PHILOX_M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
        k0 += PHILOX_W0;
        k1 += PHILOX_W1;
    }
    return c;
}

__host__ __device__ __forceinline__ uint4 philox_block(uint64_t key, uint32_t c0, uint32_t episode, uint32_t stream,
                                                       uint32_t c3) {
    return philox4x32_10(make_uint4(c0, episode, stream, c3), (uint32_t)key, (uint32_t)(key >> 32));
}

// uniform in [0,1) with 52 random bits from two 32-bit words: the bits become the mantissa of a double in [1,2)
// (two logic ops and one subtraction instead of a 64-bit integer-to-double conversion)
__host__ __device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
#ifdef __CUDA_ARCH__
    return __hiloint2double((int)(0x3FF00000u | (hi >> 12)), (int)lo) - 1.0;
#else
    uint64_t x = ((uint64_t)hi << 32) | lo;
    return (double)(x >> 12) * (1.0 / 4503599627370496.0);
#endif
}

// one alias-table draw from ONE 32-bit random word: the top log2k bits pick the bucket, the remaining bits
// (left-aligned to 32) are compared with the bucket's acceptance threshold.  Probabilities are therefore resolved
// to 2^-(32-log2k) per bucket (<= 2.4e-10 absolute for the 128-bucket Poisson(20) table) -- far below what any
// goodness-of-fit test on <= 1e12 samples can see -- and one Philox block feeds four periods.
__device__ __forceinline__ int32_t alias_draw(const uint2* __restrict__ table, int log2k, int base, uint32_t w) {
    uint32_t idx = log2k ? (w >> (32 - log2k)) : 0u;
    uint32_t frac = w << log2k;
    uint2 e = table[idx];
    return base + (int32_t)(frac < e.x ? idx : e.y);
}

__device__ __forceinline__ uint32_t pick_word(const uint4& w, int i) {
    return i == 0 ? w.x : (i == 1 ? w.y : (i == 2 ? w.z : w.w));
}

// demand for (key, episode, period t, source r) from a fixed distribution
__device__ __forceinline__ int64_t sample_fixed(const AliasDev& A, const uint2* table, uint64_t key, uint32_t episode,
                                                int t, uint32_t source) {
    if (A.kind == ORGYM_DIST_USER) {
        int idx = t;
        if (A.user_clamp) idx = t < A.user_D_len - 1 ? t : A.user_D_len - 1;
        return (idx >= 0 && idx < A.user_D_len) ? A.user_D[idx] : 0;
    }
    uint4 w = philox_block(key, (uint32_t)t >> 2, episode, STREAM_DEMAND, source);
    return alias_draw(table, A.log2k, A.base, pick_word(w, t & 3));
}

