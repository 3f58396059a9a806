// launch_util.cuh -- host-side launch helpers (no device code)
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

// Opt a kernel into `smem` bytes of dynamic shared memory.  The attribute is remembered per kernel and device and only
// ever raised, so steady-state launches skip the driver call (it showed up as microseconds per step() at small batches).
template <auto Fn>
static inline void orgym_ensure_dyn_smem(size_t smem) {
    static size_t have[64] = {0};
    int dev = -1;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || smem > have[dev]) {
        cudaFuncSetAttribute(Fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (dev >= 0 && dev < 64) have[dev] = smem;
    }
}
