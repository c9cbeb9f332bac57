// Host-only stand-in for <cuda_runtime.h>: lets g++ parse csrc/common.cuh + csrc/stencil.cuh as
// plain C++ so that (a) the compile-time tile logic (st_tile, slot / mask helpers) and (b) the
// WHOLE stencil kernels (k_apply_stencil_tma, k_apply_stencil, k_observe_stencil) can be executed
// on the CPU by the harnesses in tests/cpu_emul/.  TEST INFRASTRUCTURE ONLY - never part of the
// product.
//
// Execution model of lm_emul::launch: CTAs run one after the other; inside a CTA every CUDA thread
// is a real OS thread, so __syncthreads is a real barrier, warp shuffles exchange through a per-warp
// mailbox, atomicAdd is atomic, and the mbarrier / cp.async.bulk pair is restated with the
// hardware's phase rule (a phase completes when the pending arrivals AND the transaction byte
// count both reach zero; bytes may complete before the barrier is armed).
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>
#define LM_CPU_EMUL 1
#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __shared__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
struct double2 { double x, y; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x = 1, y = 1, z = 1; dim3() {} dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
typedef void* cudaStream_t;
static thread_local uint3 threadIdx, blockIdx, blockDim, gridDim;
inline double2 make_double2(double a, double b) { return double2{a, b}; }
inline float2 make_float2(float a, float b) { return float2{a, b}; }
template <typename T> inline T __ldg(const T* p) { return *p; }
template <typename T> inline T __ldcs(const T* p) { return *p; }
template <typename T> inline void __stcs(T* p, T v) { *p = v; }
inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }

namespace lm_emul {
struct Cta {
    unsigned nthreads;
    std::barrier<> cta_bar;
    std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
    std::vector<double> mailbox;                    // [nthreads]
    explicit Cta(unsigned nt) : nthreads(nt), cta_bar(nt), mailbox(nt) {
        for (unsigned w = 0; w < (nt + 31) / 32; ++w) {
            const unsigned lanes = (w * 32 + 32 <= nt) ? 32 : nt - w * 32;
            warp_bar.emplace_back(new std::barrier<>(lanes));
        }
    }
};
inline Cta*& cta() { static Cta* c = nullptr; return c; }
inline unsigned char* dyn_smem() { alignas(128) static unsigned char buf[232448]; return buf; }

struct MBar { long long tx = 0; int pending = 0, count = 0; unsigned phase = 0; };
inline std::mutex& mbar_mutex() { static std::mutex m; return m; }
inline std::map<const void*, MBar>& mbars() { static std::map<const void*, MBar> m; return m; }
inline void mbar_settle(MBar& b) { if (b.pending == 0 && b.tx == 0) { b.phase ^= 1u; b.pending = b.count; } }
inline long long& bulk_bytes() { static long long n = 0; return n; }

// grid.x * grid.y CTAs of `nt` threads each; blockIdx.x fastest (the hardware's launch order is
// unspecified, the kernels may not depend on it)
template <typename K, typename A>
inline void launch(K kernel, dim3 grid, unsigned nt, const A& args) {
    for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
            Cta c(nt);
            cta() = &c;
            { std::lock_guard<std::mutex> g(mbar_mutex()); mbars().clear(); }
            std::vector<std::thread> th;
            for (unsigned t = 0; t < nt; ++t)
                th.emplace_back([&, t] {
                    threadIdx = uint3{t, 0, 0}; blockIdx = uint3{bx, by, 0};
                    blockDim = uint3{nt, 1, 1}; gridDim = uint3{grid.x, grid.y, 1};
                    kernel(args);
                    c.cta_bar.arrive_and_drop();     // a thread that returned no longer takes part in __syncthreads
                });
            for (auto& t : th) t.join();
            cta() = nullptr;
        }
}
}  // namespace lm_emul

inline void __syncthreads() { if (lm_emul::cta()) lm_emul::cta()->cta_bar.arrive_and_wait(); }
inline void __syncwarp() {}
inline double __shfl_xor_sync(unsigned, double v, int o) {
    lm_emul::Cta* c = lm_emul::cta();
    if (!c) return v;
    const unsigned t = threadIdx.x, w = t >> 5;
    c->mailbox[t] = v;
    c->warp_bar[w]->arrive_and_wait();
    const double r = c->mailbox[t ^ (unsigned)o];
    c->warp_bar[w]->arrive_and_wait();
    return r;
}
inline double atomicAdd(double* p, double v) { return std::atomic_ref<double>(*p).fetch_add(v); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return std::atomic_ref<unsigned>(*p).fetch_add(v); }
inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) {
    std::atomic_ref<unsigned long long> a(*p);
    unsigned long long o = a.load();
    while (o < v && !a.compare_exchange_weak(o, v)) {}
    return o;
}
inline long long __double_as_longlong(double v) { long long r; std::memcpy(&r, &v, sizeof r); return r; }
inline void sincos(double x, double* s, double* c) { *s = std::sin(x); *c = std::cos(x); }
inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __nanosleep(unsigned) { std::this_thread::yield(); }
template <typename T> inline T __ldcv(const T* p) { return *(const volatile T*)p; }
using std::fmax; using std::fmin; using std::fabs; using std::sqrt; using std::acos; using std::ceil;

// dynamic / static shared memory of the kernels (csrc/common.cuh defines the CUDA forms)
#define LM_SMEM_DYN(name) unsigned char* name = lm_emul::dyn_smem()
#define LM_SMEM_STATIC static

namespace lm {
inline void st_release_sys(unsigned long long* p, unsigned long long v) { std::atomic_ref<unsigned long long>(*p).store(v, std::memory_order_release); }
inline unsigned long long ld_acquire_sys(const unsigned long long* p) { return std::atomic_ref<unsigned long long>(*const_cast<unsigned long long*>(p)).load(std::memory_order_acquire); }
inline void pdl_wait() {}
inline void pdl_launch_dependents() {}
inline unsigned smem_u32(const void* p) { return (unsigned)(size_t)p; }
inline void mbar_init(unsigned long long* bar, unsigned count) {
    std::lock_guard<std::mutex> g(lm_emul::mbar_mutex());
    lm_emul::MBar& b = lm_emul::mbars()[bar];
    b = lm_emul::MBar(); b.count = (int)count; b.pending = (int)count;
}
inline void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    std::lock_guard<std::mutex> g(lm_emul::mbar_mutex());
    lm_emul::MBar& b = lm_emul::mbars().at(bar);
    b.tx += bytes; b.pending -= 1; lm_emul::mbar_settle(b);
}
inline void mbar_arrive(unsigned long long* bar) {
    std::lock_guard<std::mutex> g(lm_emul::mbar_mutex());
    lm_emul::MBar& b = lm_emul::mbars().at(bar);
    b.pending -= 1; lm_emul::mbar_settle(b);
}
inline void mbar_wait(unsigned long long* bar, unsigned parity) {
    for (;;) {
        {
            std::lock_guard<std::mutex> g(lm_emul::mbar_mutex());
            if (lm_emul::mbars().at(bar).phase != parity) return;    // the phase with this parity has completed
        }
        std::this_thread::yield();
    }
}
// cp.async.bulk: 16-byte aligned addresses, size a multiple of 16 (checked: a violation is a
// hardware fault on the device)
inline void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    if ((size_t)dst % 16 || (size_t)src % 16 || bytes % 16 || bytes == 0) {
        fprintf(stderr, "EMUL FAULT: cp.async.bulk misaligned (dst %p src %p bytes %u)\n", dst, src, bytes);
        std::abort();
    }
    std::memcpy(dst, src, bytes);
    std::lock_guard<std::mutex> g(lm_emul::mbar_mutex());
    lm_emul::MBar& b = lm_emul::mbars().at(bar);
    b.tx -= bytes; lm_emul::bulk_bytes() += bytes; lm_emul::mbar_settle(b);
}
}  // namespace lm
