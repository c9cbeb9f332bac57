// The slice of the CUDA runtime API that csrc/api.cu + csrc/stencil*.cu use, restated on host
// memory for the CPU execution harness (see cuda_runtime.h in this directory).  Work "enqueued on a
// stream" executes at call time; under stream capture it is recorded as a closure and replayed by
// cudaGraphLaunch - kernel arguments are evaluated when the launch is recorded, as on the device.
// TEST INFRASTRUCTURE ONLY.
#pragma once
#include <sys/mman.h>

enum cudaError_t { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotSupported = 801 };
inline const char* cudaGetErrorString(cudaError_t e) {
    switch (e) {
    case cudaSuccess: return "no error";
    case cudaErrorInvalidValue: return "invalid argument (emulated runtime)";
    case cudaErrorMemoryAllocation: return "out of memory (emulated runtime)";
    default: return "operation not supported by the emulated runtime";
    }
}
inline cudaError_t cudaGetLastError() { return cudaSuccess; }

struct lm_emul_stream { int id; };
typedef lm_emul_stream* cudaStream_t;
struct lm_emul_event { std::chrono::steady_clock::time_point t; };
typedef lm_emul_event* cudaEvent_t;
typedef std::vector<std::function<void()>> lm_emul_graph;
typedef lm_emul_graph* cudaGraph_t;
typedef lm_emul_graph* cudaGraphExec_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal = 0, cudaStreamCaptureModeThreadLocal = 1, cudaStreamCaptureModeRelaxed = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrL2CacheSize = 38 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaIpcMemLazyEnablePeerAccess = 1 };
struct cudaIpcMemHandle_t { char reserved[64]; };

namespace lm_emul {
inline lm_emul_graph*& capturing() { static lm_emul_graph* g = nullptr; return g; }
inline void submit(std::function<void()> job) { if (capturing()) capturing()->push_back(std::move(job)); else job(); }
inline long long& live_allocs() { static long long n = 0; return n; }
// <<<grid, block, smem, stream>>>(args...) of the product sources is rewritten by build_emul_lib.py to
//     lm_emul::enqueue(grid, block, [](auto... a_) { kernel(a_...); }, args...)
template <typename K, typename... A>
inline void enqueue(dim3 grid, dim3 block, K kernel, A... args) {
    auto tup = std::make_tuple(args...);
    submit([grid, block, kernel, tup] { run_grid(grid, block, [&] { std::apply(kernel, tup); }); });
}
}  // namespace lm_emul

inline cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidValue; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int) {
    *v = (a == cudaDevAttrMultiProcessorCount) ? 148 : (a == cudaDevAttrL2CacheSize ? 126 * 1024 * 1024 : 0);
    return cudaSuccess;
}
// Device allocations END at a PROT_NONE guard page (16-byte granularity): a kernel or copy that runs
// past a buffer - silently tolerated by a real GPU more often than not - faults here.
namespace lm_emul {
struct Alloc { void* map; size_t map_bytes; size_t bytes; };
inline std::map<void*, Alloc>& allocs() { static std::map<void*, Alloc> m; return m; }
}
template <typename T> inline cudaError_t cudaMalloc(T** p, size_t bytes) {
    const size_t page = 4096, need = (bytes + 15) & ~(size_t)15;
    const size_t body = (need + page - 1) / page * page;
    void* map = mmap(nullptr, body + page, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (map == MAP_FAILED) return cudaErrorMemoryAllocation;
    mprotect((char*)map + body, page, PROT_NONE);
    char* q = (char*)map + body - need;
    std::memset(map, 0xA5, body);                   // device memory is not zeroed: make reliance on it visible
    lm_emul::allocs()[q] = lm_emul::Alloc{map, body + page, bytes};
    *p = (T*)q; ++lm_emul::live_allocs();
    return cudaSuccess;
}
inline cudaError_t cudaFree(void* p) {
    if (!p) return cudaSuccess;
    auto it = lm_emul::allocs().find(p);
    if (it == lm_emul::allocs().end()) { fprintf(stderr, "EMUL FAULT: cudaFree of a pointer cudaMalloc did not return (%p)\n", p); std::abort(); }
    munmap(it->second.map, it->second.map_bytes);
    lm_emul::allocs().erase(it); --lm_emul::live_allocs();
    return cudaSuccess;
}
template <typename T> inline cudaError_t cudaMallocHost(T** p, size_t bytes) { *p = (T*)std::malloc(bytes ? bytes : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) {
    lm_emul::submit([d, s, n] { if (n) std::memmove(d, s, n); });
    return cudaSuccess;
}
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t = nullptr) {
    lm_emul::submit([=] { for (size_t r = 0; r < height; ++r) std::memmove((char*)d + r * dpitch, (const char*)s + r * spitch, width); });
    return cudaSuccess;
}
inline cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) {
    lm_emul::submit([d, v, n] { if (n) std::memset(d, v, n); });
    return cudaSuccess;
}
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new lm_emul_stream{1}; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return lm_emul::capturing() ? cudaErrorNotSupported : cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new lm_emul_event{std::chrono::steady_clock::now()}; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return cudaSuccess;
}
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int bytes) { return bytes <= 227 * 1024 ? cudaSuccess : cudaErrorInvalidValue; }
inline cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) {
    if (lm_emul::capturing()) return cudaErrorInvalidValue;
    lm_emul::capturing() = new lm_emul_graph;
    return cudaSuccess;
}
inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) { *g = lm_emul::capturing(); lm_emul::capturing() = nullptr; return *g ? cudaSuccess : cudaErrorInvalidValue; }
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t g, unsigned long long = 0) { *e = new lm_emul_graph(*g); return cudaSuccess; }
inline cudaError_t cudaGraphDestroy(cudaGraph_t g) { delete g; return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t) { for (auto& job : *e) lm_emul::submit(job); return cudaSuccess; }
// programmatic launches (stencil_inst.cuh): same execution, the attribute is a scheduling hint
enum cudaLaunchAttributeID { cudaLaunchAttributeProgrammaticStreamSerialization = 4 };
struct cudaLaunchAttributeValue { int programmaticStreamSerializationAllowed; };
struct cudaLaunchAttribute { cudaLaunchAttributeID id; cudaLaunchAttributeValue val; };
struct cudaLaunchConfig_t { dim3 gridDim, blockDim; size_t dynamicSmemBytes; cudaStream_t stream; cudaLaunchAttribute* attrs; unsigned numAttrs; };
template <typename K, typename... A> inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t* cfg, K kernel, A... arg) {
    lm_emul::enqueue(cfg->gridDim, cfg->blockDim, kernel, arg...);
    return cudaSuccess;
}
// multi-GPU plumbing: not available in the single-process harness
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*) { return cudaErrorNotSupported; }
inline cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaErrorNotSupported; }
