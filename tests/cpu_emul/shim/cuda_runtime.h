// Host-only stand-in for <cuda_runtime.h>: lets g++ compile the product's CUDA sources as plain
// C++ so that they can be EXECUTED ON THE CPU by the harnesses in tests/cpu_emul/:
//   (a) the compile-time tile logic (st_tile, slot / mask helpers)           stencil_emul.cpp
//   (b) whole kernels against reference loops / the oracle                    stencil_kernel_emul.cpp, kernels_emul.cpp
//   (c) the whole library - C ABI, host orchestration, every kernel           build_emul_lib.py
// TEST INFRASTRUCTURE ONLY - never part of the product, never loaded by it.
//
// Execution model: CTAs run one after the other; inside a CTA every CUDA thread is a fiber with its
// own stack, resumed round-robin, so __syncthreads is a real barrier (threads that returned drop
// out of it), warp shuffles and the m8n8k4 DMMA exchange through a per-thread mailbox, a barrier
// nobody can complete is reported as a deadlock, and the mbarrier / cp.async.bulk pair is restated
// with the hardware's phase rule (a phase completes when the pending arrivals AND the transaction byte count
// both reach zero; bytes may complete before the barrier is armed).  The CUDA runtime calls the
// library makes are restated on host memory (cudaMalloc = aligned malloc, streams execute at call
// time, stream capture records closures that cudaGraphLaunch replays).
#pragma once
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <tuple>
#include <ucontext.h>
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
inline uint3 threadIdx, blockIdx, blockDim, gridDim;
inline double2 make_double2(double a, double b) { return double2{a, b}; }
inline float2 make_float2(float a, float b) { return float2{a, b}; }
inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
template <typename T> inline T __ldg(const T* p) { return *p; }
template <typename T> inline T __ldcs(const T* p) { return *p; }
template <typename T> inline T __ldcv(const T* p) { return *(const volatile T*)p; }
template <typename T> inline void __stcs(T* p, T v) { *p = v; }
inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }

namespace lm_emul {
// One CTA at a time; its CUDA threads are FIBERS (ucontext) of the calling OS thread, resumed
// round-robin.  A thread runs until it returns or has to wait (__syncthreads, a warp exchange, an
// mbarrier phase), then yields to the next one - no OS synchronisation, deterministic order, and a
// round in which nobody can make progress is reported as a deadlock instead of hanging.
struct Bar { unsigned active = 0, arrived = 0; unsigned long gen = 0; };
struct Cta {
    Bar cta_bar;
    Bar warp_bar[32];
    double mailbox[1024][2];
    bool progress = false;                          // somebody ran past a wait / finished during this round
};
inline Cta& cta() { static Cta c; return c; }
inline bool& in_kernel() { static bool b = false; return b; }
inline unsigned& linear_tid() { static unsigned t = 0; return t; }    // threadIdx linearised, x fastest
inline unsigned char* dyn_smem() { alignas(128) static unsigned char buf[232448]; return buf; }

struct MBar { long long tx = 0; int pending = 0, count = 0; unsigned phase = 0; };
inline std::map<const void*, MBar>& mbars() { static std::map<const void*, MBar> m; return m; }
inline void mbar_settle(MBar& b) { if (b.pending == 0 && b.tx == 0) { b.phase ^= 1u; b.pending = b.count; } }
inline long long& bulk_bytes() { static long long n = 0; return n; }
inline long long& cta_count() { static long long n = 0; return n; }

// context switch: on x86-64 a hand-written callee-saved-register swap (glibc's swapcontext makes a
// signal-mask system call per switch, which dominates the run time of sync-free kernels); ucontext elsewhere
#if defined(__x86_64__)
extern "C" void lm_emul_switch(void** save_sp, void* load_sp);
asm(".text\n.weak lm_emul_switch\n.type lm_emul_switch,@function\nlm_emul_switch:\n"
    "pushq %rbp\npushq %rbx\npushq %r12\npushq %r13\npushq %r14\npushq %r15\n"
    "movq %rsp, (%rdi)\nmovq %rsi, %rsp\n"
    "popq %r15\npopq %r14\npopq %r13\npopq %r12\npopq %rbx\npopq %rbp\nret\n"
    ".size lm_emul_switch, .-lm_emul_switch\n");
struct Context { void* sp = nullptr; };
inline void ctx_switch(Context& from, Context& to) { lm_emul_switch(&from.sp, to.sp); }
inline void ctx_make(Context& c, void* stack, size_t size, void (*fn)()) {
    uintptr_t top = ((uintptr_t)stack + size) & ~(uintptr_t)15;
    void** p = (void**)top;
    *--p = nullptr;                 // fake return address of fn (it never returns)
    *--p = (void*)fn;               // popped by the `ret` of the first switch: rsp = top - 8 at entry, as after a call
    for (int i = 0; i < 6; ++i) *--p = nullptr;
    c.sp = p;
}
#else
struct Context { ucontext_t uc; };
inline void ctx_switch(Context& from, Context& to) { swapcontext(&from.uc, &to.uc); }
inline void ctx_make(Context& c, void* stack, size_t size, void (*fn)()) {
    getcontext(&c.uc);
    c.uc.uc_stack.ss_sp = stack; c.uc.uc_stack.ss_size = size; c.uc.uc_link = nullptr;
    makecontext(&c.uc, fn, 0);
}
#endif
struct Fiber { Context ctx; void* stack = nullptr; bool done = true; };
struct Sched {
    static constexpr size_t kStack = 256 * 1024;
    Context main_ctx;
    std::vector<Fiber> fibers;
    const std::function<void()>* body = nullptr;
    unsigned cur = 0;
    uint3 bdim{1, 1, 1};
    static void entry() {
        Sched& s = sched();
        (*s.body)();
        Fiber& f = s.fibers[s.cur];
        f.done = true;
        Cta& c = cta();
        c.progress = true;
        // a thread that returned no longer takes part in the barriers
        for (Bar* b : {&c.cta_bar, &c.warp_bar[s.cur >> 5]}) {
            if (b->active) --b->active;
            if (b->active && b->arrived >= b->active) { b->arrived = 0; ++b->gen; }
        }
        ctx_switch(f.ctx, s.main_ctx);
    }
    static Sched& sched() { static Sched s; return s; }
    void yield() { ctx_switch(fibers[cur].ctx, main_ctx); }
    void run_cta(uint3 bd, const std::function<void()>& f) {
        const unsigned n = bd.x * bd.y * bd.z;
        if (fibers.size() < n) fibers.resize(n);
        Cta& c = cta();
        c.cta_bar = Bar{n, 0, 0};
        for (unsigned w = 0; w < 32; ++w) c.warp_bar[w] = Bar{w * 32 < n ? ((w * 32 + 32 <= n) ? 32 : n - w * 32) : 0, 0, 0};
        mbars().clear();
        body = &f; bdim = bd;
        for (unsigned t = 0; t < n; ++t) {
            Fiber& fb = fibers[t];
            if (!fb.stack) fb.stack = std::malloc(kStack);
            ctx_make(fb.ctx, fb.stack, kStack, &Sched::entry);
            fb.done = false;
        }
        in_kernel() = true;
        unsigned left = n;
        while (left) {
            c.progress = false;
            for (unsigned t = 0; t < n; ++t) {
                if (fibers[t].done) continue;
                cur = t; linear_tid() = t;
                threadIdx = uint3{t % bd.x, (t / bd.x) % bd.y, t / (bd.x * bd.y)};
                ctx_switch(main_ctx, fibers[t].ctx);
                if (fibers[t].done) --left;
            }
            if (left && !c.progress) {
                fprintf(stderr, "EMUL FAULT: deadlock - %u thread(s) of CTA (%u,%u,%u) wait on a barrier that cannot complete\n", left, blockIdx.x, blockIdx.y, blockIdx.z);
                std::abort();
            }
        }
        in_kernel() = false;
        ++cta_count();
    }
};
// wait until `gen` of the barrier moves on; the last arriver releases everybody
inline void bar_wait(Bar& b) {
    Cta& c = cta();
    c.progress = true;                              // an arrival is progress; only re-checking waiters are not
    if (++b.arrived >= b.active) { b.arrived = 0; ++b.gen; return; }
    const unsigned long g = b.gen;
    while (b.gen == g) Sched::sched().yield();
    c.progress = true;
}

// grid.x * grid.y * grid.z CTAs, blockIdx.x fastest (the hardware's order is unspecified)
inline void run_grid(dim3 grid, dim3 block, const std::function<void()>& thread_body) {
    const unsigned nt = block.x * block.y * block.z;
    if (nt == 0 || nt > 1024) { fprintf(stderr, "EMUL FAULT: %u threads per CTA\n", nt); std::abort(); }
    static std::mutex one_grid_at_a_time;
    std::lock_guard<std::mutex> guard(one_grid_at_a_time);
    blockDim = uint3{block.x, block.y, block.z}; gridDim = uint3{grid.x, grid.y, grid.z};
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                blockIdx = uint3{bx, by, bz};
                Sched::sched().run_cta(uint3{block.x, block.y, block.z}, thread_body);
            }
}
template <typename K, typename A>
inline void launch(K kernel, dim3 grid, unsigned nt, const A& args) { run_grid(grid, dim3(nt), [&] { kernel(args); }); }
}  // namespace lm_emul

inline void __syncthreads() { if (lm_emul::in_kernel()) lm_emul::bar_wait(lm_emul::cta().cta_bar); }
inline void __syncwarp() {}
inline double __shfl_xor_sync(unsigned, double v, int o) {
    if (!lm_emul::in_kernel()) return v;
    lm_emul::Cta& c = lm_emul::cta();
    const unsigned t = lm_emul::linear_tid(), w = t >> 5;
    c.mailbox[t][0] = v;
    lm_emul::bar_wait(c.warp_bar[w]);
    const double r = c.mailbox[t ^ (unsigned)o][0];
    lm_emul::bar_wait(c.warp_bar[w]);
    return r;
}
inline double atomicAdd(double* p, double v) { return std::atomic_ref<double>(*p).fetch_add(v); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return std::atomic_ref<unsigned>(*p).fetch_add(v); }
inline unsigned atomicOr(unsigned* p, unsigned v) { return std::atomic_ref<unsigned>(*p).fetch_or(v); }
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
inline void __nanosleep(unsigned) { if (lm_emul::in_kernel()) lm_emul::Sched::sched().yield(); }
using std::fmax; using std::fmin; using std::fabs; using std::sqrt; using std::acos; using std::ceil;

// dynamic / static shared memory of the kernels (csrc/common.cuh defines the CUDA forms)
#define LM_SMEM_DYN(name) unsigned char* name = lm_emul::dyn_smem()
#define LM_SMEM_STATIC static
#define LM_GRID_CONSTANT
// tensor map of the emulated TMA: what cuTensorMapEncodeTiled is given (FLOAT64 units)
struct CUtensorMap { const void* base; unsigned long long dims[3]; unsigned long long strides[2]; unsigned box[3]; int valid; };

namespace lm {
inline void st_release_sys(unsigned long long* p, unsigned long long v) { std::atomic_ref<unsigned long long>(*p).store(v, std::memory_order_release); }
inline unsigned long long ld_acquire_sys(const unsigned long long* p) { return std::atomic_ref<unsigned long long>(*const_cast<unsigned long long*>(p)).load(std::memory_order_acquire); }
inline void pdl_wait() {}
inline void pdl_launch_dependents() {}
inline unsigned smem_u32(const void* p) { return (unsigned)(size_t)p; }
inline void mbar_init(unsigned long long* bar, unsigned count) {
    lm_emul::MBar& b = lm_emul::mbars()[bar];
    b = lm_emul::MBar(); b.count = (int)count; b.pending = (int)count;
}
inline void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    lm_emul::MBar& b = lm_emul::mbars().at(bar);
    b.tx += bytes; b.pending -= 1; lm_emul::mbar_settle(b);
}
inline void mbar_arrive(unsigned long long* bar) {
    lm_emul::MBar& b = lm_emul::mbars().at(bar);
    b.pending -= 1; lm_emul::mbar_settle(b);
}
inline void mbar_wait(unsigned long long* bar, unsigned parity) {
    while (lm_emul::mbars().at(bar).phase == parity) lm_emul::Sched::sched().yield();   // until the phase with this parity has completed
    lm_emul::cta().progress = true;
}
// cp.async.bulk: 16-byte aligned addresses, size a multiple of 16 (checked: a violation is a
// hardware fault on the device)
inline void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    if ((size_t)dst % 16 || (size_t)src % 16 || bytes % 16 || bytes == 0) {
        fprintf(stderr, "EMUL FAULT: cp.async.bulk misaligned (dst %p src %p bytes %u)\n", dst, src, bytes);
        std::abort();
    }
    std::memcpy(dst, src, bytes);
    lm_emul::MBar& b = lm_emul::mbars().at(bar);
    b.tx -= bytes; lm_emul::bulk_bytes() += bytes; lm_emul::mbar_settle(b);
    lm_emul::cta().progress = true;
}
// cp.async.bulk.tensor.3d (tile mode, FLOAT64 elements): the whole box lands densely in shared memory,
// innermost dimension first; elements outside the tensor are zero-filled and counted in the bytes
inline void tma_tensor3d_g2s(void* dst, const void* tmap, int c0, int c1, int c2, unsigned long long* bar) {
    const CUtensorMap& t = *(const CUtensorMap*)tmap;
    if (!t.valid || (size_t)dst % 128) { fprintf(stderr, "EMUL FAULT: tensor copy with an invalid map / misaligned destination\n"); std::abort(); }
    double* d = (double*)dst;
    for (unsigned k2 = 0; k2 < t.box[2]; ++k2)
        for (unsigned k1 = 0; k1 < t.box[1]; ++k1)
            for (unsigned k0 = 0; k0 < t.box[0]; ++k0) {
                const long long i0 = (long long)c0 + k0, i1 = (long long)c1 + k1, i2 = (long long)c2 + k2;
                const bool in = i0 >= 0 && i1 >= 0 && i2 >= 0 && i0 < (long long)t.dims[0] && i1 < (long long)t.dims[1] && i2 < (long long)t.dims[2];
                *d++ = in ? *(const double*)((const char*)t.base + (size_t)i2 * t.strides[1] + (size_t)i1 * t.strides[0] + (size_t)i0 * 8) : 0.0;
            }
    const long long bytes = 8LL * t.box[0] * t.box[1] * t.box[2];
    lm_emul::MBar& b = lm_emul::mbars().at(bar);
    b.tx -= bytes; lm_emul::bulk_bytes() += bytes; lm_emul::mbar_settle(b);
    lm_emul::cta().progress = true;
}
inline void tma_tensor3d_prefetch_l2(const void*, int, int, int) {}      // a cache hint: nothing to emulate
// mma.sync.aligned.m8n8k4.row.col.f64: lane l holds A[l>>2][l&3], B[l&3][l>>2], C[l>>2][2(l&3) + {0,1}]
inline void dmma(double& d0, double& d1, double a, double b) {
    lm_emul::Cta& c = lm_emul::cta();
    const unsigned t = lm_emul::linear_tid(), base = t & ~31u, l = t & 31u, w = t >> 5;
    c.mailbox[t][0] = a; c.mailbox[t][1] = b;
    lm_emul::bar_wait(c.warp_bar[w]);
    const unsigned i = l >> 2, j0 = 2 * (l & 3);
    for (unsigned k = 0; k < 4; ++k) {
        const double aik = c.mailbox[base + i * 4 + k][0];
        d0 += aik * c.mailbox[base + j0 * 4 + k][1];
        d1 += aik * c.mailbox[base + (j0 + 1) * 4 + k][1];
    }
    lm_emul::bar_wait(c.warp_bar[w]);
}
}  // namespace lm

#include "cuda_runtime_api_emul.h"
