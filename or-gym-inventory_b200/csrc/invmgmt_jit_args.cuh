// invmgmt_jit_args.cuh -- launch arguments of the run-time specialised serial-env rollout kernels; shared verbatim by
// the host (invmgmt.cu) and the generated source (invmgmt_jit.cu embeds this text for NVRTC).  Plain fixed-width types.
#pragma once
struct InvJitArgs {
    int64_t N, env_offset;
    uint64_t seed;
    uint32_t episode;
    int32_t target[16];   // base-stock levels (integers; the kernel is only used when they are)
    const uint2* table;   // demand alias table (global memory; staged in shared memory by the kernel)
    double* ep_return;    // [N] or null
    int64_t* stats;       // [N,4] or null
    int32_t* stats32;     // [N,4] or null
    double* partials;     // [gridDim.x][8] or null
};
