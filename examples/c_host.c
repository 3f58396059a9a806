/* Minimal C host for liborgym_b200.so: the C ABI of include/orgym_b200.h without Python or torch.
 *
 * Runs N instances of the reference's default InvManagementBacklogEnv (inventory_management.py:19-167 defaults) for one
 * 30-period episode through reset/step with a constant order, then the same episode count through the fused rollout
 * with the on-device base-stock policy, and prints the mean episode returns.
 *
 *   gcc -std=c99 -O2 -I include -I /usr/local/cuda/include examples/c_host.c \
 *       -L or-gym-inventory_b200/csrc -lorgym_b200 -L /usr/local/cuda/lib64 -lcudart -lm -o c_host
 *   LD_LIBRARY_PATH=or-gym-inventory_b200/csrc ./c_host
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "orgym_b200.h"

#define CHECK(call)                                                              \
    do {                                                                         \
        int rc_ = (call);                                                        \
        if (rc_ != 0) {                                                          \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, orgym_last_error()); \
            return 1;                                                            \
        }                                                                        \
    } while (0)
#define CUDA(call)                                                               \
    do {                                                                         \
        cudaError_t e_ = (call);                                                 \
        if (e_ != cudaSuccess) {                                                 \
            fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_));          \
            return 1;                                                            \
        }                                                                        \
    } while (0)

int main(void) {
    enum { M = 4, NS = 3, T = 30 };
    const int64_t N = 4096;
    /* the reference's defaults; prices/costs are float32-rounded like its np.array(..., dtype=np.float32) (:89-92) */
    int64_t I0[NS] = {100, 150, 200}, cap[NS] = {100, 200, 230}, L[NS] = {1, 5, 10};
    double up[M] = {20.0, 15.0, 10.0, 7.0}, uc[M] = {15.0, 10.0, 7.0, 5.0}; /* [p]+r[:-1] and r with p=20, r=[15,10,7,5] */
    double kc[M] = {(float)0.10, (float)0.075, (float)0.05, (float)0.025}, hc[M] = {(float)0.15, (float)0.10, (float)0.05, 0.0};
    orgym_invmgmt_config_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.num_stages = M;
    cfg.periods = T;
    cfg.backlog = 1;
    cfg.init_inv = I0;
    cfg.capacity = cap;
    cfg.lead_time = L;
    cfg.unit_price = up;
    cfg.unit_cost = uc;
    cfg.demand_cost = kc;
    cfg.holding_cost = hc;
    cfg.alpha = 0.97;
    cfg.dist.kind = 1; /* Poisson */
    cfg.dist.p0 = 20.0;

    if (orgym_device_count() < 1) {
        fprintf(stderr, "no CUDA device: %s\n", orgym_last_error());
        return 2;
    }
    orgym_handle_t h;
    CHECK(orgym_invmgmt_create(&cfg, N, 0, &h));
    const int od = orgym_invmgmt_obs_dim(h);
    void *state, *obs, *act, *rew, *term, *trunc, *ret, *summary;
    CUDA(cudaMalloc(&state, (size_t)orgym_invmgmt_state_bytes(h)));
    CUDA(cudaMalloc(&obs, sizeof(int64_t) * N * od));
    CUDA(cudaMalloc(&act, sizeof(int64_t) * N * NS));
    CUDA(cudaMalloc(&rew, sizeof(double) * N));
    CUDA(cudaMalloc(&term, N));
    CUDA(cudaMalloc(&trunc, N));
    CUDA(cudaMalloc(&ret, sizeof(double) * N));
    CUDA(cudaMalloc(&summary, sizeof(double) * 8));

    /* --- Env.reset / Env.step loop with a constant order of 20 units per stage --- */
    int64_t* a_host = (int64_t*)malloc(sizeof(int64_t) * N * NS);
    double* r_host = (double*)malloc(sizeof(double) * N);
    for (int64_t i = 0; i < N * NS; i++) a_host[i] = 20;
    CUDA(cudaMemcpy(act, a_host, sizeof(int64_t) * N * NS, cudaMemcpyHostToDevice));
    CHECK(orgym_invmgmt_reset(h, state, 1, 4000, 0, NULL, (int64_t*)obs, NULL));
    double total = 0.0;
    for (int t = 0; t < T; t++) {
        CHECK(orgym_invmgmt_step(h, state, act, 0, NULL, ORGYM_AUTORESET_DISABLED, (int64_t*)obs, (double*)rew, (uint8_t*)term, (uint8_t*)trunc, NULL, NULL));
        CUDA(cudaMemcpy(r_host, rew, sizeof(double) * N, cudaMemcpyDeviceToHost));
        for (int64_t i = 0; i < N; i++) total += r_host[i];
    }
    printf("step API,  constant order 20: mean episode return %.4f over %lld instances\n", total / (double)N, (long long)N);

    /* --- evaluate_agent + BaseStockAgent fused into one launch --- */
    orgym_invmgmt_rollout_in_t in;
    orgym_invmgmt_rollout_out_t out;
    memset(&in, 0, sizeof(in));
    memset(&out, 0, sizeof(out));
    in.policy = ORGYM_POLICY_BASE_STOCK;
    in.param[0] = 1.0;   /* safety factor */
    in.param[1] = 20.0;  /* mu */
    out.ep_return_dev = (double*)ret;
    out.summary_dev = (double*)summary;
    CHECK(orgym_invmgmt_rollout(h, 4000, 0, 0, &in, &out, NULL));
    double s[8];
    CUDA(cudaMemcpy(s, summary, sizeof(s), cudaMemcpyDeviceToHost));
    printf("rollout,   base-stock SF=1.0:  mean episode return %.4f over %.0f instances (std %.2f)\n", s[1] / s[0], s[0],
           sqrt(s[2] / s[0] - (s[1] / s[0]) * (s[1] / s[0])));
    uint32_t err = 0;
    CHECK(orgym_errors(h, &err, 0, NULL));
    printf("device error flags: %u\n", err);
    CHECK(orgym_invmgmt_destroy(h));
    free(a_host);
    free(r_host);
    return err != 0;
}
