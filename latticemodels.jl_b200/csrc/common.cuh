// Shared device helpers of the sm_100a kernels: complex arithmetic on double2 / float2, 128-bit
// lane elements, cache-hinted loads and stores.
#pragma once
#ifdef __CUDACC_RTC__
// run-time compilation of a stencil kernel for a lattice pattern that is not among the compiled ones (csrc/stencil_rtc.cpp):
// NVRTC has the CUDA built-ins but no host headers; the tensor map is an opaque 128-byte kernel parameter
struct alignas(64) CUtensorMap_st { unsigned long long opaque[16]; };
typedef CUtensorMap_st CUtensorMap;
#else
#include <cuda_runtime.h>
#include <stdint.h>
#ifndef LM_CPU_EMUL
#include <cuda.h>               // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#endif
#endif

namespace lm {

template <typename T> struct cx2;
template <> struct cx2<double> { using type = double2; };
template <> struct cx2<float>  { using type = float2; };

template <typename T2> __device__ __forceinline__ T2 cmake(double re, double im) {
    T2 r; r.x = (decltype(r.x))re; r.y = (decltype(r.y))im; return r;
}
template <typename T2> __device__ __forceinline__ void cfma(T2& acc, const T2 a, const T2 b) {
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
template <typename T2> __device__ __forceinline__ T2 cmul(const T2 a, const T2 b) {
    T2 r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r;
}

// read-only (ld.global.nc) L1-allocating loads for the gathered Psi rows (re-used by the
// neighbouring rows of the same CTA); evict-first streaming loads / stores for operands that
// are touched exactly once per pass.  Intrinsics (not volatile asm) so that ptxas is free to
// batch the independent gathers of a row ahead of the FMA chain (memory-level parallelism).
__device__ __forceinline__ double2 ld_ro(const double2* p) { return __ldg(p); }
__device__ __forceinline__ float2 ld_ro(const float2* p) { return __ldg(p); }
__device__ __forceinline__ double2 ld_stream(const double2* p) { return __ldcs(p); }
__device__ __forceinline__ float2 ld_stream(const float2* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(double2* p, double2 v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(float2* p, float2 v) { __stcs(p, v); }
// "pinned" variants: volatile asm keeps the load where it is written (ptxas otherwise sinks the
// element-wise operand loads below the gather/FMA section to save registers, which serialises
// two HBM latencies per thread).
__device__ __forceinline__ double2 ld_stream_pin(const double2* p) {
    double2 r;
    asm volatile("ld.global.cs.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ float2 ld_stream_pin(const float2* p) {
    float2 r;
    asm volatile("ld.global.cs.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p) : "memory");
    return r;
}

// What ONE LANE moves per 128-bit access: one complex128, or TWO adjacent complex64 columns
// (so that the complex64 mode keeps full-width L1 / HBM transactions).
template <typename T> struct pack;
template <> struct pack<double> { using E = double2; static constexpr int EC = 1; };
template <> struct pack<float>  { using E = float4;  static constexpr int EC = 2; };
__device__ __forceinline__ float4 ld_ro(const float4* p) { return __ldg(p); }
__device__ __forceinline__ float4 ld_stream(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float4* p, float4 v) { __stcs(p, v); }
__device__ __forceinline__ void pzero(double2& a) { a.x = 0; a.y = 0; }
__device__ __forceinline__ void pzero(float4& a) { a.x = 0; a.y = 0; a.z = 0; a.w = 0; }
__device__ __forceinline__ void pfma(double2& acc, const double2 v, const double2 x) { cfma(acc, v, x); }
__device__ __forceinline__ void pfma(float4& acc, const float2 v, const float4 x) {
    acc.x = fmaf(v.x, x.x, acc.x); acc.x = fmaf(-v.y, x.y, acc.x);
    acc.y = fmaf(v.x, x.y, acc.y); acc.y = fmaf(v.y, x.x, acc.y);
    acc.z = fmaf(v.x, x.z, acc.z); acc.z = fmaf(-v.y, x.w, acc.z);
    acc.w = fmaf(v.x, x.w, acc.w); acc.w = fmaf(v.y, x.z, acc.w);
}

// acc += conj(v) * x  (the Hermitian partner of a bond shares the value load, stencil.cuh st_tile_herm)
__device__ __forceinline__ void pfma_conj(double2& acc, const double2 v, const double2 x) {
    acc.x = fma(v.x, x.x, acc.x); acc.x = fma(v.y, x.y, acc.x);
    acc.y = fma(v.x, x.y, acc.y); acc.y = fma(-v.y, x.x, acc.y);
}
__device__ __forceinline__ void pfma_conj(float4& acc, const float2 v, const float4 x) {
    acc.x = fmaf(v.x, x.x, acc.x); acc.x = fmaf(v.y, x.y, acc.x);
    acc.y = fmaf(v.x, x.y, acc.y); acc.y = fmaf(-v.y, x.x, acc.y);
    acc.z = fmaf(v.x, x.z, acc.z); acc.z = fmaf(v.y, x.w, acc.z);
    acc.w = fmaf(v.x, x.w, acc.w); acc.w = fmaf(-v.y, x.z, acc.w);
}

// acc += s * x for a purely REAL value s, acc += (i s) * x for a purely IMAGINARY value i s: two FMAs per complex
// element instead of four, and the value is a 64-bit (32-bit) scalar (stencil.cuh, the "real / imaginary" value class)
__device__ __forceinline__ void pfma_re(double2& acc, const double s, const double2 x) {
    acc.x = fma(s, x.x, acc.x); acc.y = fma(s, x.y, acc.y);
}
__device__ __forceinline__ void pfma_im(double2& acc, const double s, const double2 x) {
    acc.x = fma(-s, x.y, acc.x); acc.y = fma(s, x.x, acc.y);
}
__device__ __forceinline__ void pfma_re(float4& acc, const float s, const float4 x) {
    acc.x = fmaf(s, x.x, acc.x); acc.y = fmaf(s, x.y, acc.y); acc.z = fmaf(s, x.z, acc.z); acc.w = fmaf(s, x.w, acc.w);
}
__device__ __forceinline__ void pfma_im(float4& acc, const float s, const float4 x) {
    acc.x = fmaf(-s, x.y, acc.x); acc.y = fmaf(s, x.x, acc.y); acc.z = fmaf(-s, x.w, acc.z); acc.w = fmaf(s, x.z, acc.w);
}

// ---- shared-memory declarations of the staged kernels ----
// (macros so that the CPU execution harness, tests/cpu_emul/, can substitute host storage;
//  under nvcc they expand to exactly the usual CUDA declarations)
#ifndef LM_CPU_EMUL
#define LM_GRID_CONSTANT __grid_constant__
#define LM_SMEM_DYN(name) extern __shared__ __align__(128) unsigned char name[]
#define LM_SMEM_STATIC __shared__
#endif

// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) ----
// (the CPU execution harness restates these five primitives in tests/cpu_emul/shim/cuda_runtime.h)
#ifndef LM_CPU_EMUL
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "LM_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra LM_DONE_%=;\n\t"
        "bra LM_WAIT_%=;\n\t"
        "LM_DONE_%=:\n\t}"
        :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// 1-D bulk async copy global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// 3-D tiled tensor-map copy global -> shared (TMA, SASS UTMALDG): one instruction moves the whole
// box; coordinates are element indices, innermost first; out-of-range elements are zero-filled and
// still counted in the transaction bytes
__device__ __forceinline__ void tma_tensor3d_g2s(void* dst, const void* tmap, int c0, int c1, int c2, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
// L2 prefetch of the same box (no shared-memory destination, no barrier): issued for the patch a LATER CTA will
// stage, so that its tensor-map copy is served by L2 instead of HBM
__device__ __forceinline__ void tma_tensor3d_prefetch_l2(const void* tmap, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
                 :: "l"(tmap), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// programmatic dependent launch (griddepcontrol): no-ops for a grid launched without the attribute
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif  // LM_CPU_EMUL

}  // namespace lm
