// Host-only stand-in for <cuda_runtime.h>: lets g++ parse csrc/common.cuh + csrc/stencil.cuh as
// plain C++ so that the compile-time tile logic (st_tile, slot / mask helpers) can be executed on
// the CPU by tests/cpu_emul/stencil_emul.cpp.  TEST INFRASTRUCTURE ONLY - never part of the product.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
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
static uint3 threadIdx, blockIdx, blockDim, gridDim;
inline double2 make_double2(double a, double b) { return double2{a, b}; }
inline float2 make_float2(float a, float b) { return float2{a, b}; }
template <typename T> inline T __ldg(const T* p) { return *p; }
template <typename T> inline T __ldcs(const T* p) { return *p; }
template <typename T> inline void __stcs(T* p, T v) { *p = v; }
inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
inline void __syncthreads() {}
inline void __syncwarp() {}
inline double __shfl_xor_sync(unsigned, double v, int) { return v; }
inline double atomicAdd(double* p, double v) { double o = *p; *p += v; return o; }
