// jit.cu -- NVRTC front-end (see jit.cuh)
#include "jit.cuh"

#include <dlfcn.h>

#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "jit_embed.inc"
const char* orgym_jit_device_rng_src() { return k_device_rng_src; }
const char* orgym_jit_net_args_src() { return k_net_args_src; }
const char* orgym_jit_inv_args_src() { return k_inv_args_src; }

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
    int (*Version)(int*, int*) = nullptr;
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
    SYM(Version, "nvrtcVersion")
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

// ---- cubin cache: in-process map + files under $ORGYM_JIT_CACHE (default ~/.cache/orgym_b200) -----------------------
// Key: 64-bit FNV-1a hash of the generated source (which embeds every constant of the config), its length, the NVRTC
// version actually loaded and the compile options.  A cache file starts with a header repeating hash / source length /
// cubin length, checked on load; the directory is created 0700 and ignored unless it belongs to the current user and is
// not writable by anyone else (a cubin is code that will run on the GPU: never load one somebody else could have planted).
static const char* const k_nvrtc_opts[] = {"--gpu-architecture=sm_100a", "--fmad=false", "-lineinfo", "--std=c++17",
                                           "-default-device"};
static const int k_nvrtc_nopts = 5;
static std::mutex g_cache_mu;
static std::map<std::string, std::vector<char>> g_cubin_cache;

static uint64_t fnv1a(const std::string& src, uint64_t h = 1469598103934665603ULL) {
    for (unsigned char c : src) {
        h ^= c;
        h *= 1099511628211ULL;
    }
    return h;
}
static std::string cache_key(const std::string& src) {
    std::call_once(g_once, load_nvrtc);
    int vmaj = 0, vmin = 0;
    if (g_nvrtc.ok) g_nvrtc.Version(&vmaj, &vmin);
    std::string opts;
    for (int i = 0; i < k_nvrtc_nopts; i++) opts += std::string(k_nvrtc_opts[i]) + " ";
    char buf[96];
    snprintf(buf, sizeof(buf), "%016llx_%zu_nvrtc%d.%d_%08x", (unsigned long long)fnv1a(src), src.size(), vmaj, vmin,
             (unsigned)(fnv1a(opts) & 0xffffffffu));
    return buf;
}
struct CacheHeader {
    char magic[8];  // "ORGYMJIT"
    uint64_t src_hash, src_len, cubin_len;
};
static std::string cache_dir() {
    const char* d = getenv("ORGYM_JIT_CACHE");
    if (d && d[0]) return d[0] == '0' && d[1] == 0 ? std::string() : std::string(d);
    const char* home = getenv("HOME");
    if (!home || !home[0]) return std::string();
    return std::string(home) + "/.cache/orgym_b200";
}
// the directory exists, belongs to us and nobody else may write into it
static bool cache_dir_trusted(const std::string& dir) {
    struct stat st;
    if (stat(dir.c_str(), &st) != 0 || !S_ISDIR(st.st_mode)) return false;
    return st.st_uid == geteuid() && (st.st_mode & (S_IWGRP | S_IWOTH)) == 0;
}
static bool cache_load(const std::string& key, const std::string& src, std::vector<char>* out) {
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        auto it = g_cubin_cache.find(key);
        if (it != g_cubin_cache.end()) {
            *out = it->second;
            return true;
        }
    }
    std::string dir = cache_dir();
    if (dir.empty() || !cache_dir_trusted(dir)) return false;
    FILE* f = fopen((dir + "/" + key + ".cubin").c_str(), "rb");
    if (!f) return false;
    CacheHeader hd;
    bool ok = fread(&hd, 1, sizeof(hd), f) == sizeof(hd) && memcmp(hd.magic, "ORGYMJIT", 8) == 0 &&
              hd.src_hash == fnv1a(src) && hd.src_len == src.size() && hd.cubin_len > 0 && hd.cubin_len < (1u << 28);
    if (ok) {
        out->resize((size_t)hd.cubin_len);
        ok = fread(out->data(), 1, out->size(), f) == out->size() && fgetc(f) == EOF;
    }
    fclose(f);
    if (ok) {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        g_cubin_cache[key] = *out;
    }
    return ok;
}
static void cache_store(const std::string& key, const std::string& src, const std::vector<char>& cubin) {
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        g_cubin_cache[key] = cubin;
    }
    std::string dir = cache_dir();
    if (dir.empty()) return;
    std::string parent = dir.substr(0, dir.find_last_of('/'));
    mkdir(parent.c_str(), 0700);
    mkdir(dir.c_str(), 0700);
    if (!cache_dir_trusted(dir)) return;
    std::string tmp = dir + "/" + key + ".tmp" + std::to_string((long)getpid());
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return;
    CacheHeader hd;
    memcpy(hd.magic, "ORGYMJIT", 8);
    hd.src_hash = fnv1a(src);
    hd.src_len = src.size();
    hd.cubin_len = cubin.size();
    bool ok = fwrite(&hd, 1, sizeof(hd), f) == sizeof(hd) && fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
    fclose(f);
    if (ok)
        rename(tmp.c_str(), (dir + "/" + key + ".cubin").c_str());  // atomic publish
    else
        unlink(tmp.c_str());
}

static int nvrtc_compile(const std::string& src, std::vector<char>* cubin_out, std::string* err);

static int compile_to_cubin(const std::string& src, std::vector<char>* cubin_out, std::string* err) {
    const std::string key = cache_key(src);
    if (cache_load(key, src, cubin_out)) return 0;
    int rc = nvrtc_compile(src, cubin_out, err);
    if (rc == 0) cache_store(key, src, *cubin_out);
    return rc;
}

static int nvrtc_compile(const std::string& src, std::vector<char>* cubin_out, std::string* err) {
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
    int rc = g_nvrtc.CompileProgram(prog, k_nvrtc_nopts, k_nvrtc_opts);
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
