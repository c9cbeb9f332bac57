// Run-time specialisation of the register-tiled stencil kernels (stencil.cuh) on a DETECTED lattice pattern.
//
// The kernels are templates over the pattern (a type with the sparsity mask and the value classes as constexpr members);
// ten patterns are compiled into the library (stencil.cu).  Any other pattern of at most four rows per unit cell that
// couples adjacent cells - kagome with third neighbours, Kane-Mele with a spin-mixing term, a two-orbital model on the
// honeycomb lattice, ... - gets its own instantiation here: the same headers (embedded in the library as text,
// gen_rtc_headers.inc) are compiled by NVRTC for sm_100a with the detected mask, loaded through the driver API and
// launched with the same arguments.  Kernels are compiled lazily, one per (precision, propagator term form), and the
// cubins are cached on disk (LM_RTC_CACHE, default ~/.cache/lm_b200).  NVRTC (libnvrtc.so.12) is opened with dlopen:
// if it is missing, or LM_STENCIL_RTC=0, such lattices keep the ELL kernels - nothing fails.
#include "stencil.cuh"

#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace lm {

struct RtcKernel { void* mod = nullptr; void* fn = nullptr; bool tried = false; };
struct RtcPattern {
    StencilDesc desc;
    std::string name;
    int device = 0;                 // the modules live in this device's primary context (one process per GPU: always the same one)
    RtcKernel apply[2][4];          // [complex64][MODE]
    RtcKernel obs[2];
};
static std::vector<RtcPattern*> g_rtc;
static std::string g_rtc_err;
const char* stencil_rtc_error() { return g_rtc_err.c_str(); }

// shapes by rows per cell: the defaults of the compiled patterns (stencil.cu variants 7 / 2 / 19, stencil_inst.cuh ObsShape)
void stencil_rtc_apply_shape(int rc, int* t1, int* t2, int* w1, int* w2) {
    *w1 = 2; *w2 = 2;
    if (rc == 1) { *t1 = 4; *t2 = 4; } else if (rc == 2) { *t1 = 4; *t2 = 2; } else { *t1 = 2; *t2 = 2; }
}
static int count_width(int rc, const st_mask_t& m) {
    int w = 1;
    for (int a = 0; a < rc; ++a) {
        int s = 0;
        for (int o = 0; o < 9; ++o) for (int b = 0; b < rc; ++b) if (st_get(m, o * rc * rc + a * rc + b)) ++s;
        if (s > w) w = s;
    }
    return w;
}
int stencil_rtc_nfwd(int rc, const st_mask_t& m) {
    int w = 1;
    for (int a = 0; a < rc; ++a) {
        int s = 0;
        for (int o = 4; o < 9; ++o) for (int b = 0; b < rc; ++b) if (st_get(m, o * rc * rc + a * rc + b) && (o > 4 || b > a)) ++s;
        if (s > w) w = s;
    }
    return w;
}
// observables kernel: T1 x T2 cells per thread such that a thread folds at most 64 partial sums; false if even one cell does not fit
bool stencil_rtc_obs_shape(int rc, int nf, int* t1, int* t2, int* w1, int* w2) {
    *w1 = 4; *w2 = 2;
    const int per_cell = rc * (1 + 2 * nf);
    if (per_cell > 64) return false;
    if (rc == 1 && 4 * per_cell <= 64) { *t1 = 2; *t2 = 2; return true; }
    *t1 = 1; *t2 = (2 * per_cell <= 64) ? 2 : 1;
    // one cell per thread: 4 x 4-cell patches (512 threads) for four rows per cell and for the widest three-row patterns
    // (measured: Kane-Mele + spin mixing 2.69 -> 2.30 ms, kagome t1 t2 t3 6.48 -> 5.39 ms; kagome NN + NNN loses, 3.37 -> 4.09 ms)
    if (*t2 == 1 && (rc == 4 || (rc == 3 && per_cell > 48))) *w2 = 4;
    return true;
}

// gen_rtc_headers.inc is written by build.py; a bare `nvcc -c stencil_rtc.cu` without it still compiles (run-time specialisation off)
#if defined(LM_CPU_EMUL)
#define LM_RTC_OFF 1
#elif defined(__has_include)
#if !__has_include("gen_rtc_headers.inc")
#define LM_RTC_OFF 1
#endif
#endif
#ifndef LM_RTC_OFF
#include "gen_rtc_headers.inc"      // g_rtc_common_cuh, g_rtc_stencil_cuh: the text of common.cuh / stencil.cuh (written by build.py)

// ---- NVRTC through dlopen ----
typedef void* nvrtcProgram_t;
struct Nvrtc {
    void* lib = nullptr;
    int (*CreateProgram)(nvrtcProgram_t*, const char*, const char*, int, const char* const*, const char* const*);
    int (*DestroyProgram)(nvrtcProgram_t*);
    int (*AddNameExpression)(nvrtcProgram_t, const char*);
    int (*CompileProgram)(nvrtcProgram_t, int, const char* const*);
    int (*GetProgramLogSize)(nvrtcProgram_t, size_t*);
    int (*GetProgramLog)(nvrtcProgram_t, char*);
    int (*GetLoweredName)(nvrtcProgram_t, const char*, const char**);
    int (*GetCUBINSize)(nvrtcProgram_t, size_t*);
    int (*GetCUBIN)(nvrtcProgram_t, char*);
    bool ok = false;
};
static Nvrtc& nvrtc() {
    static Nvrtc n = [] {
        Nvrtc r;
        const char* env = getenv("LM_NVRTC_LIB");
        const char* cands[] = {env, "libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so"};
        for (const char* c : cands) { if (!c || !*c) continue; r.lib = dlopen(c, RTLD_NOW | RTLD_LOCAL); if (r.lib) break; }
        if (!r.lib) return r;
        bool all = true;
        auto sym = [&](const char* s) { void* p = dlsym(r.lib, s); if (!p) all = false; return p; };
        r.CreateProgram = (decltype(r.CreateProgram))sym("nvrtcCreateProgram");
        r.DestroyProgram = (decltype(r.DestroyProgram))sym("nvrtcDestroyProgram");
        r.AddNameExpression = (decltype(r.AddNameExpression))sym("nvrtcAddNameExpression");
        r.CompileProgram = (decltype(r.CompileProgram))sym("nvrtcCompileProgram");
        r.GetProgramLogSize = (decltype(r.GetProgramLogSize))sym("nvrtcGetProgramLogSize");
        r.GetProgramLog = (decltype(r.GetProgramLog))sym("nvrtcGetProgramLog");
        r.GetLoweredName = (decltype(r.GetLoweredName))sym("nvrtcGetLoweredName");
        r.GetCUBINSize = (decltype(r.GetCUBINSize))sym("nvrtcGetCUBINSize");
        r.GetCUBIN = (decltype(r.GetCUBIN))sym("nvrtcGetCUBIN");
        r.ok = all;
        return r;
    }();
    return n;
}

// ---- driver API entry points (no link dependency on libcuda) ----
struct Driver {
    CUresult (*ModuleLoadData)(CUmodule*, const void*);
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*);
    CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int);
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**);
    bool ok = false;
};
static Driver& driver() {
    static Driver d = [] {
        Driver r;
        bool all = true;
        auto get = [&](const char* s) {
            void* fn = nullptr; cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint(s, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) { all = false; fn = nullptr; }
            return fn;
        };
        r.ModuleLoadData = (decltype(r.ModuleLoadData))get("cuModuleLoadData");
        r.ModuleGetFunction = (decltype(r.ModuleGetFunction))get("cuModuleGetFunction");
        r.FuncSetAttribute = (decltype(r.FuncSetAttribute))get("cuFuncSetAttribute");
        r.LaunchKernel = (decltype(r.LaunchKernel))get("cuLaunchKernel");
        r.ok = all;
        return r;
    }();
    return d;
}

static unsigned long long fnv1a(unsigned long long h, const char* p, size_t n) {
    for (size_t i = 0; i < n; ++i) { h ^= (unsigned char)p[i]; h *= 1099511628211ull; }
    return h;
}
static std::string cache_dir() {
    const char* env = getenv("LM_RTC_CACHE");
    std::string d;
    if (env && *env) d = env;
    else {
        const char* home = getenv("HOME");
        if (!home || !*home) return std::string();
        d = std::string(home) + "/.cache";
        mkdir(d.c_str(), 0755);
        d += "/lm_b200";
    }
    mkdir(d.c_str(), 0755);
    return d;
}

// the cubin of one kernel instantiation `expr` for the pattern: from the disk cache, or compiled now by NVRTC (needs no device)
static bool compile_cubin(const StencilDesc& d, const std::string& expr, std::vector<char>* cubin_out, std::string* lowered_out, bool use_cache = true) {
    char pat[640];
    snprintf(pat, sizeof(pat),
             "#include \"stencil.cuh\"\nnamespace lm { struct PatRT { static constexpr int rc = %d; static constexpr st_mask_t mask = {{0x%llxull, 0x%llxull, 0x%llxull, 0x%llxull}}, "
             "imag = {{0x%llxull, 0x%llxull, 0x%llxull, 0x%llxull}}; }; }\n",
             d.rc, d.mask.w[0], d.mask.w[1], d.mask.w[2], d.mask.w[3], d.imag.w[0], d.imag.w[1], d.imag.w[2], d.imag.w[3]);
    const std::string src = pat;
    unsigned long long key = 1469598103934665603ull;
    key = fnv1a(key, g_rtc_common_cuh, strlen(g_rtc_common_cuh));
    key = fnv1a(key, g_rtc_stencil_cuh, strlen(g_rtc_stencil_cuh));
    key = fnv1a(key, src.data(), src.size());
    key = fnv1a(key, expr.data(), expr.size());
    const std::string dir = use_cache ? cache_dir() : std::string();
    char fname[64];
    snprintf(fname, sizeof(fname), "/st_%016llx.cubin", key);
    std::vector<char>& cubin = *cubin_out;
    std::string& lowered = *lowered_out;
    cubin.clear(); lowered.clear();
    if (!dir.empty()) {
        if (FILE* f = fopen((dir + fname).c_str(), "rb")) {
            unsigned long long nl = 0, nc = 0;
            if (fread(&nl, 8, 1, f) == 1 && fread(&nc, 8, 1, f) == 1 && nl < 4096 && nc < (1ull << 30)) {
                lowered.resize(nl); cubin.resize(nc);
                if (fread(&lowered[0], 1, nl, f) != nl || fread(cubin.data(), 1, nc, f) != nc) { lowered.clear(); cubin.clear(); }
            }
            fclose(f);
        }
    }
    if (!cubin.empty()) return true;
    Nvrtc& rt = nvrtc();
    if (!rt.ok) { g_rtc_err = "libnvrtc.so.12 not found (LM_NVRTC_LIB)"; return false; }
    nvrtcProgram_t prog = nullptr;
    const char* hdr_src[] = {g_rtc_stencil_cuh, g_rtc_common_cuh};
    const char* hdr_name[] = {"stencil.cuh", "common.cuh"};
    if (rt.CreateProgram(&prog, src.c_str(), "lm_stencil_rtc.cu", 2, hdr_src, hdr_name) != 0) { g_rtc_err = "nvrtcCreateProgram failed"; return false; }
    rt.AddNameExpression(prog, expr.c_str());
    // -default-device: the generic lambdas of the tile bodies carry no execution-space annotation
    const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device"};
    const int rc = rt.CompileProgram(prog, 4, opts);
    if (rc != 0) {
        size_t n = 0; rt.GetProgramLogSize(prog, &n);
        std::string log(n, ' ');
        if (n) rt.GetProgramLog(prog, &log[0]);
        g_rtc_err = "NVRTC compilation of " + expr + " failed:\n" + log.substr(0, 2000);
        rt.DestroyProgram(&prog);
        return false;
    }
    const char* low = nullptr;
    size_t n = 0;
    if (rt.GetLoweredName(prog, expr.c_str(), &low) != 0 || !low || rt.GetCUBINSize(prog, &n) != 0 || n == 0) {
        g_rtc_err = "NVRTC produced no cubin for " + expr; rt.DestroyProgram(&prog); return false;
    }
    lowered = low;
    cubin.resize(n);
    rt.GetCUBIN(prog, cubin.data());
    rt.DestroyProgram(&prog);
    if (!dir.empty()) {
        const std::string tmp = dir + fname + "." + std::to_string((long long)getpid());
        if (FILE* f = fopen(tmp.c_str(), "wb")) {
            const unsigned long long nl = lowered.size(), nc = cubin.size();
            const bool w = fwrite(&nl, 8, 1, f) == 1 && fwrite(&nc, 8, 1, f) == 1 && fwrite(lowered.data(), 1, nl, f) == nl && fwrite(cubin.data(), 1, nc, f) == nc;
            fclose(f);
            if (w) rename(tmp.c_str(), (dir + fname).c_str()); else unlink(tmp.c_str());
        }
    }
    return true;
}
// compile / fetch one kernel instantiation, load it, set its shared-memory limit
static bool build_kernel(const RtcPattern& p, const std::string& expr, size_t smem, RtcKernel* out) {
    out->tried = true;
    Driver& drv = driver();
    if (!drv.ok) { g_rtc_err = "driver entry points unavailable"; return false; }
    std::vector<char> cubin;
    std::string lowered;
    if (!compile_cubin(p.desc, expr, &cubin, &lowered)) return false;
    CUmodule mod = nullptr; CUfunction fn = nullptr;
    if (drv.ModuleLoadData(&mod, cubin.data()) != CUDA_SUCCESS) { g_rtc_err = "cuModuleLoadData failed for " + expr; return false; }
    if (drv.ModuleGetFunction(&fn, mod, lowered.c_str()) != CUDA_SUCCESS) { g_rtc_err = "cuModuleGetFunction failed for " + expr; return false; }
    if (smem > 48 * 1024 && drv.FuncSetAttribute(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem) != CUDA_SUCCESS) {
        g_rtc_err = "cuFuncSetAttribute(shared memory) failed for " + expr; return false;
    }
    out->mod = mod; out->fn = fn;
    return true;
}
// NVRTC alone, no device: the size of the cubin of the apply kernel (mode 0 .. 3) or of the observables kernel (mode < 0) for a pattern, -1 on failure
long long stencil_rtc_compile_only(int rc, const st_mask_t& mask, const st_mask_t& imag, bool c64, int mode) {
    StencilDesc d{rc, mask, imag, count_width(rc, mask), "probe"};
    int t1, t2, w1, w2;
    char expr[256];
    if (mode >= 0) {
        stencil_rtc_apply_shape(rc, &t1, &t2, &w1, &w2);
        snprintf(expr, sizeof(expr), "lm::k_apply_stencil_tma<%s, %d, lm::PatRT, %d, %d, %d, %d, 1, %d>", c64 ? "float" : "double", rc, t1, t2, w1, w2, mode);
    } else {
        if (!stencil_rtc_obs_shape(rc, stencil_rtc_nfwd(rc, mask), &t1, &t2, &w1, &w2)) { g_rtc_err = "too many forward entries per cell for the observables kernel"; return -1; }
        snprintf(expr, sizeof(expr), "lm::k_observe_stencil<%s, %d, lm::PatRT, %d, %d, %d, %d>", c64 ? "float" : "double", rc, t1, t2, w1, w2);
    }
    std::vector<char> cubin; std::string lowered;
    if (!compile_cubin(d, expr, &cubin, &lowered, false)) return -1;
    return (long long)cubin.size();
}
static bool rtc_enabled() {
    const char* v = getenv("LM_STENCIL_RTC");
    return !(v && atoi(v) == 0);
}
bool stencil_rtc_available() { return rtc_enabled() && nvrtc().ok && driver().ok; }

int stencil_rtc_launch(int id, bool c64, int mode, const StencilArgs& a_in, const CUtensorMap& tmx, dim3 grid, cudaStream_t s) {
    RtcPattern& p = *g_rtc[id - LM_ST_RTC_BASE];
    if (mode < 0 || mode > 3) return -1;
    int t1, t2, w1, w2;
    stencil_rtc_apply_shape(p.desc.rc, &t1, &t2, &w1, &w2);
    const int P1 = t1 * w1, P2 = t2 * w2, sw = stencil_stride(id, c64);
    const size_t smem = (size_t)(P1 + 2) * (P2 + 2) * p.desc.rc * 32 * 16 + (size_t)P1 * P2 * p.desc.rc * sw * (c64 ? 8 : 16);
    RtcKernel& k = p.apply[c64 ? 1 : 0][mode];
    if (!k.tried) {
        char expr[256];
        snprintf(expr, sizeof(expr), "lm::k_apply_stencil_tma<%s, %d, lm::PatRT, %d, %d, %d, %d, 1, %d>", c64 ? "float" : "double", p.desc.rc, t1, t2, w1, w2, mode);
        build_kernel(p, expr, smem, &k);
    }
    if (!k.fn) return -1;
    StencilArgs a = a_in;
    a.pdl = 0;
    CUtensorMap tm = tmx;
    void* params[] = {&a, &tm};
    return driver().LaunchKernel((CUfunction)k.fn, grid.x, grid.y, 1, 32 * w1 * w2, 1, 1, (unsigned)smem, (CUstream)s, params, nullptr) == CUDA_SUCCESS ? 0 : -2;
}
int stencil_rtc_observe(int id, bool c64, const StencilObsArgs& a_in, const CUtensorMap& tmx, unsigned grid, cudaStream_t s) {
    RtcPattern& p = *g_rtc[id - LM_ST_RTC_BASE];
    int t1, t2, w1, w2;
    if (!stencil_rtc_obs_shape(p.desc.rc, stencil_rtc_nfwd(p.desc.rc, p.desc.mask), &t1, &t2, &w1, &w2)) return -1;
    const size_t smem = (size_t)3 * (w1 * t1 + 1) * (w2 * t2 + 2) * p.desc.rc * 32 * 16;       // st_obs_smem (ST_OBS_STAGES = 3)
    static_assert(ST_OBS_STAGES == 3, "stencil_rtc_observe restates st_obs_smem");
    RtcKernel& k = p.obs[c64 ? 1 : 0];
    if (!k.tried) {
        char expr[256];
        snprintf(expr, sizeof(expr), "lm::k_observe_stencil<%s, %d, lm::PatRT, %d, %d, %d, %d>", c64 ? "float" : "double", p.desc.rc, t1, t2, w1, w2);
        build_kernel(p, expr, smem, &k);
    }
    if (!k.fn) return -1;
    StencilObsArgs a = a_in;
    CUtensorMap tm = tmx;
    void* params[] = {&a, &tm};
    return driver().LaunchKernel((CUfunction)k.fn, grid, 1, 1, 32 * w1 * w2, 1, 1, (unsigned)smem, (CUstream)s, params, nullptr) == CUDA_SUCCESS ? 0 : -2;
}
// compile the kernels a propagation step and a frame need (plain SpMM, product-form factor, observables) now instead of at first use
int stencil_rtc_warm(int id, bool c64) {
    RtcPattern& p = *g_rtc[id - LM_ST_RTC_BASE];
    int t1, t2, w1, w2;
    stencil_rtc_apply_shape(p.desc.rc, &t1, &t2, &w1, &w2);
    const int P1 = t1 * w1, P2 = t2 * w2, sw = stencil_stride(id, c64);
    const size_t smem = (size_t)(P1 + 2) * (P2 + 2) * p.desc.rc * 32 * 16 + (size_t)P1 * P2 * p.desc.rc * sw * (c64 ? 8 : 16);
    if (smem > 227 * 1024) { g_rtc_err = "patch does not fit shared memory"; return -1; }
    for (int mode : {3, 0}) {
        RtcKernel& k = p.apply[c64 ? 1 : 0][mode];
        if (k.tried) { if (!k.fn) return -1; continue; }
        char expr[256];
        snprintf(expr, sizeof(expr), "lm::k_apply_stencil_tma<%s, %d, lm::PatRT, %d, %d, %d, %d, 1, %d>", c64 ? "float" : "double", p.desc.rc, t1, t2, w1, w2, mode);
        if (!build_kernel(p, expr, smem, &k)) return -1;
    }
    return 0;
}
#else   // CPU execution harness / no embedded headers: no run-time compilation
long long stencil_rtc_compile_only(int, const st_mask_t&, const st_mask_t&, bool, int) { return -1; }
bool stencil_rtc_available() { return false; }
int stencil_rtc_launch(int, bool, int, const StencilArgs&, const CUtensorMap&, dim3, cudaStream_t) { return -1; }
int stencil_rtc_observe(int, bool, const StencilObsArgs&, const CUtensorMap&, unsigned, cudaStream_t) { return -1; }
int stencil_rtc_warm(int, bool) { return -1; }
#endif

// register a detected pattern (or find it again); returns its id >= LM_ST_RTC_BASE, -1 if run-time compilation is not available
int stencil_rtc_register(int rc, const st_mask_t& mask, const st_mask_t& imag) {
    if (rc < 1 || rc > 4 || !stencil_rtc_available()) return -1;
    int dev = 0;
#ifndef LM_CPU_EMUL
    cudaGetDevice(&dev);
#endif
    for (size_t i = 0; i < g_rtc.size(); ++i) {
        const StencilDesc& d = g_rtc[i]->desc;
        if (g_rtc[i]->device == dev && d.rc == rc && !memcmp(&d.mask, &mask, sizeof(mask)) && !memcmp(&d.imag, &imag, sizeof(imag))) return LM_ST_RTC_BASE + (int)i;
    }
    RtcPattern* p = new RtcPattern;
    p->device = dev;
    char nm[96];
    snprintf(nm, sizeof(nm), "rtc-rc%d-%llx:%llx:%llx", rc, mask.w[2], mask.w[1], mask.w[0]);
    p->name = nm;
    p->desc = StencilDesc{rc, mask, imag, count_width(rc, mask), p->name.c_str()};
    g_rtc.push_back(p);
    return LM_ST_RTC_BASE + (int)g_rtc.size() - 1;
}
const StencilDesc* stencil_rtc_desc(int id) {
    const int i = id - LM_ST_RTC_BASE;
    return (i >= 0 && i < (int)g_rtc.size()) ? &g_rtc[i]->desc : nullptr;
}

}  // namespace lm

// debug / test entry point: NVRTC compilation of a pattern's kernel without a device (tests/test_host_cpu.py); cubin size or -1
extern "C" long long lm_dbg_rtc_compile(int rc, const unsigned long long* mask4, const unsigned long long* imag4, int c64, int mode) {
    lm::st_mask_t m = {{mask4[0], mask4[1], mask4[2], mask4[3]}}, im = {{imag4[0], imag4[1], imag4[2], imag4[3]}};
    return lm::stencil_rtc_compile_only(rc, m, im, c64 != 0, mode);
}
extern "C" const char* lm_dbg_rtc_error(void) { return lm::stencil_rtc_error(); }
