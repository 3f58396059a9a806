// netinv_args.cuh -- launch arguments of the network-env kernels; shared verbatim by the ahead-of-time kernel
// (netinv.cu) and the run-time specialised kernel (netinv_jit.cu embeds this text for NVRTC).  Plain fixed-width types.
#pragma once
#ifndef ORGYM_AUTORESET_NEXT_STEP
#define ORGYM_AUTORESET_DISABLED 0
#define ORGYM_AUTORESET_NEXT_STEP 1
#define ORGYM_AUTORESET_SAME_STEP 2
#define ORGYM_ERR_STEP_PAST_END 1u
#define ORGYM_NET_POLICY_ACTIONS 0
#define ORGYM_NET_POLICY_CONSTANT 1
#endif

// element (env e, column k) of a two-dimensional info tensor with `dim` columns
#define NET_IIDX(A, e, dim, k) ((A).info_ld ? (long long)(k) * (A).info_ld + (e) : (e) * (long long)(dim) + (k))

struct NetSimArgs {
    int64_t N, env_offset;
    int rollout;  // 0 = one period from / to state (STEP), 1 = fused episode (ROLLOUT)
    void* state;  // STEP: live state; ROLLOUT: scratch for the rings
    uint64_t seed;
    uint32_t episode;
    int policy;
    const float* actions;
    int64_t a_se, a_st;
    const double* demand;
    int64_t d_se, d_st;
    int autoreset;
    // STEP outputs
    float* obs;
    double* reward;
    uint8_t* terminated;
    uint8_t* truncated;
    double* info_demand;
    double* info_sales;
    double* info_profit;
    double* info_profit_total;
    int64_t info_ld;  // 0: info tensors row-major [N][dim]; > 0: column-major [dim][info_ld]
    float* final_obs;
    uint32_t* err;
    int use_tile;  // STEP: stage the observation block in shared memory and store it coalesced
    // ROLLOUT outputs
    double* ep_return;
    double* stats;
    double* reward_traj;
    double* final_X;
    double* final_Y;
    double* final_U;
    double* partials;
};

