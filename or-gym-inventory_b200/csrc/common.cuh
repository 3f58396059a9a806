// common.cuh -- shared host/device utilities of liborgym_b200.so (sm_100a only).
//
//  * error plumbing for the C ABI (thread-local last error, CUDA checks)
//  * Philox4x32-10 counter-based generator and the stream/counter conventions
//  * alias-table sampling for the fixed demand distributions (per-instance Poisson means: poisson_mu.cuh)
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
