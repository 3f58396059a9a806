// jit.cuh -- run-time kernel specialisation: CUDA source generated from an env config is compiled for sm_100a with
// NVRTC (loaded with dlopen, so the library has no hard dependency on it) and launched through the CUDA runtime's
// library API.  If NVRTC is missing or the compile fails the caller falls back to the generic ahead-of-time kernel.
#pragma once
#include <cuda_runtime.h>

#include <string>

struct JitKernel {
    cudaLibrary_t lib = nullptr;
    cudaKernel_t fn = nullptr;
};

// Compiles `src` (a complete translation unit) and fetches the extern "C" kernel `name`.  Returns 0 on success;
// on failure returns nonzero and puts the reason (including the NVRTC log) into *err.
int orgym_jit_compile(const std::string& src, const char* name, JitKernel* out, std::string* err);
void orgym_jit_release(JitKernel* k);
// compile only (no GPU needed): used by the code-generation self test
int orgym_jit_compile_only(const std::string& src, std::string* err);
// text of device_rng.cuh, embedded at build time (jit_embed.inc)
const char* orgym_jit_device_rng_src();
// text of netinv_args.cuh
const char* orgym_jit_net_args_src();
// text of invmgmt_jit_args.cuh
const char* orgym_jit_inv_args_src();
