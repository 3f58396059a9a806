// jit.cu -- NVRTC front-end (see jit.cuh)
#include "jit.cuh"

#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "jit_embed.inc"
const char* orgym_jit_device_rng_src() { return k_device_rng_src; }
const char* orgym_jit_net_args_src() { return k_net_args_src; }

namespace {
typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
    void* h = nullptr;
    int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    int (*DestroyProgram)(nvrtcProgram*) = nullptr;
    bool ok = false;
};
Nvrtc g_nvrtc;
std::once_flag g_once;

void load_nvrtc() {
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                           "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* n : names) {
        g_nvrtc.h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (g_nvrtc.h) break;
    }
    if (!g_nvrtc.h) return;
#define SYM(field, name)                                                   \
    g_nvrtc.field = (decltype(g_nvrtc.field))dlsym(g_nvrtc.h, name);       \
    if (!g_nvrtc.field) return;
    SYM(CreateProgram, "nvrtcCreateProgram")
    SYM(CompileProgram, "nvrtcCompileProgram")
    SYM(GetCUBINSize, "nvrtcGetCUBINSize")
    SYM(GetCUBIN, "nvrtcGetCUBIN")
    SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
    SYM(GetProgramLog, "nvrtcGetProgramLog")
    SYM(DestroyProgram, "nvrtcDestroyProgram")
#undef SYM
    g_nvrtc.ok = true;
}
}  // namespace

static int compile_to_cubin(const std::string& src, std::vector<char>* cubin_out, std::string* err);

int orgym_jit_compile_only(const std::string& src, std::string* err) {
    std::vector<char> cubin;
    return compile_to_cubin(src, &cubin, err);
}

int orgym_jit_compile(const std::string& src, const char* name, JitKernel* out, std::string* err) {
    std::vector<char> cubin;
    int rc = compile_to_cubin(src, &cubin, err);
    if (rc != 0) return rc;
    cudaError_t e = cudaLibraryLoadData(&out->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e != cudaSuccess) {
        *err = std::string("cudaLibraryLoadData failed: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return 4;
    }
    e = cudaLibraryGetKernel(&out->fn, out->lib, name);
    if (e != cudaSuccess) {
        *err = std::string("cudaLibraryGetKernel failed: ") + cudaGetErrorString(e);
        cudaGetLastError();
        cudaLibraryUnload(out->lib);
        out->lib = nullptr;
        return 5;
    }
    return 0;
}

static int compile_to_cubin(const std::string& src, std::vector<char>* cubin_out, std::string* err) {
    std::call_once(g_once, load_nvrtc);
    if (!g_nvrtc.ok) {
        *err = "NVRTC (libnvrtc.so.12) could not be loaded";
        return 1;
    }
    nvrtcProgram prog = nullptr;
    if (g_nvrtc.CreateProgram(&prog, src.c_str(), "orgym_jit.cu", 0, nullptr, nullptr) != 0) {
        *err = "nvrtcCreateProgram failed";
        return 2;
    }
    const char* opts[] = {"--gpu-architecture=sm_100a", "--fmad=false", "-lineinfo", "--std=c++17", "-default-device"};
    int rc = g_nvrtc.CompileProgram(prog, 5, opts);
    if (rc != 0) {
        size_t n = 0;
        g_nvrtc.GetProgramLogSize(prog, &n);
        std::vector<char> log(n + 1, 0);
        if (n) g_nvrtc.GetProgramLog(prog, log.data());
        *err = std::string("NVRTC compile failed: ") + log.data();
        g_nvrtc.DestroyProgram(&prog);
        if (getenv("ORGYM_JIT_DUMP")) {
            FILE* f = fopen(getenv("ORGYM_JIT_DUMP"), "w");
            if (f) {
                fputs(src.c_str(), f);
                fclose(f);
            }
        }
        return 3;
    }
    size_t nb = 0;
    g_nvrtc.GetCUBINSize(prog, &nb);
    cubin_out->resize(nb);
    g_nvrtc.GetCUBIN(prog, cubin_out->data());
    g_nvrtc.DestroyProgram(&prog);
    if (getenv("ORGYM_JIT_DUMP")) {
        FILE* f = fopen(getenv("ORGYM_JIT_DUMP"), "w");
        if (f) {
            fputs(src.c_str(), f);
            fclose(f);
        }
    }
    return 0;
}

void orgym_jit_release(JitKernel* k) {
    if (k->lib) cudaLibraryUnload(k->lib);
    k->lib = nullptr;
    k->fn = nullptr;
}
