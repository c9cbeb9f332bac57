// Launch plumbing of the stencil kernels, included by the per-pattern translation units
// stencil_k<id>.cu (the fully unrolled kernels are large: one pattern per unit, built in parallel).
#pragma once
#include "stencil.cuh"

namespace lm {

template <typename T, int RC, typename MK, int T1, int T2, int W1, int W2, int CPT, int MODE, int STAGED>
static int launch_one(const StencilArgs& a, const CUtensorMap& tmx, dim3 grid, cudaStream_t s) {
    if constexpr (STAGED == 2) {
        static_assert(CPT == 1, "the streaming kernel handles one lane element per thread");
        constexpr size_t smem = st_stream_smem<T, RC, MK, T1, T2, W1, W2>();
        static bool configured = false;
        if (!configured) {
            if (cudaFuncSetAttribute(k_apply_stencil_stream<T, RC, MK, T1, T2, W1, W2, MODE>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -2;
            configured = true;
        }
        k_apply_stencil_stream<T, RC, MK, T1, T2, W1, W2, MODE><<<grid, 32 * W1 * W2, smem, s>>>(a);
        return 0;
    } else if constexpr (STAGED == 0) {
        k_apply_stencil<T, RC, MK, T1, T2, W1, W2, CPT, MODE><<<grid, 32 * W1 * W2, 0, s>>>(a);
        return 0;
    } else {
    constexpr size_t smem = st_tma_smem<T, RC, MK, T1, T2, W1, W2, CPT>();
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(k_apply_stencil_tma<T, RC, MK, T1, T2, W1, W2, CPT, MODE>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -2;
        configured = true;
    }
    if (a.pdl) {
        // programmatic stream serialization: this grid may start while the previous one of the
        // stream drains (the kernel waits with griddepcontrol.wait before touching memory)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid; cfg.blockDim = dim3(32 * W1 * W2); cfg.dynamicSmemBytes = smem; cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, k_apply_stencil_tma<T, RC, MK, T1, T2, W1, W2, CPT, MODE>, a, tmx) == cudaSuccess ? 0 : -2;
    }
    k_apply_stencil_tma<T, RC, MK, T1, T2, W1, W2, CPT, MODE><<<grid, 32 * W1 * W2, smem, s>>>(a, tmx);
    return 0;
    }
}
// STAGED is a compile-time family switch so that only the requested family is instantiated
template <typename T, int RC, typename MK, int T1, int T2, int W1, int W2, int CPT, int STAGED>
static int launch_modes(int mode, const StencilArgs& a, const CUtensorMap& tmx, dim3 grid, cudaStream_t s) {
    switch (mode) {
    case 0: return launch_one<T, RC, MK, T1, T2, W1, W2, CPT, 0, STAGED>(a, tmx, grid, s);
    case 3: return launch_one<T, RC, MK, T1, T2, W1, W2, CPT, 3, STAGED>(a, tmx, grid, s);
#ifndef LM_STENCIL_FEWMODES
    case 1: return launch_one<T, RC, MK, T1, T2, W1, W2, CPT, 1, STAGED>(a, tmx, grid, s);
    case 2: return launch_one<T, RC, MK, T1, T2, W1, W2, CPT, 2, STAGED>(a, tmx, grid, s);
#endif
    default: return -1;
    }
}
template <int RC, typename MK, int T1, int T2, int W1, int W2, int CPT, int STAGED>
static int launch_prec(bool c64, int mode, const StencilArgs& a, const CUtensorMap& tmx, dim3 grid, cudaStream_t s) {
    if (!c64) return launch_modes<double, RC, MK, T1, T2, W1, W2, CPT, STAGED>(mode, a, tmx, grid, s);
#ifndef LM_STENCIL_NOC64
    return launch_modes<float, RC, MK, T1, T2, W1, W2, CPT, STAGED>(mode, a, tmx, grid, s);
#else
    return -1;
#endif
}

#define LM_ST_V(v, T1, T2, W1, W2, CPT, ST) case v: return launch_prec<RC, MK, T1, T2, W1, W2, CPT, ST>(c64, mode, a, tmx, grid, s);
// shape experiments: complex128, plain SpMM and product-form factor only (keeps the build short)
template <int RC, typename MK, int T1, int T2, int W1, int W2, int CPT>
static int launch_try(bool c64, int mode, const StencilArgs& a, const CUtensorMap& tmx, dim3 grid, cudaStream_t s) {
    if (c64) return -1;
    if (mode == 0) return launch_one<double, RC, MK, T1, T2, W1, W2, CPT, 0, 1>(a, tmx, grid, s);
    if (mode == 3) return launch_one<double, RC, MK, T1, T2, W1, W2, CPT, 3, 1>(a, tmx, grid, s);
    return -1;
}
#define LM_ST_X(v, T1, T2, W1, W2, CPT) case v: return launch_try<RC, MK, T1, T2, W1, W2, CPT>(c64, mode, a, tmx, grid, s);
template <int RC, typename MK>
static int launch_var(int variant, bool c64, int mode, const StencilArgs& a, const CUtensorMap& tmx, dim3 grid, cudaStream_t s) {
    if constexpr (RC == 1) {
        switch (variant) {
            LM_ST_V(7, 4, 4, 2, 2, 1, 1)
#ifdef LM_STENCIL_SHAPES
            LM_ST_X(16, 4, 4, 1, 2, 1) LM_ST_X(17, 4, 4, 2, 1, 1) LM_ST_X(18, 4, 4, 1, 1, 1)
#endif
#ifdef LM_STENCIL_EXPLORE
            LM_ST_V(8, 4, 4, 2, 2, 1, 0) LM_ST_V(9, 4, 2, 2, 4, 2, 1) LM_ST_V(3, 4, 2, 2, 4, 1, 1) LM_ST_V(10, 2, 2, 4, 2, 1, 2) LM_ST_V(11, 4, 2, 2, 2, 1, 2) LM_ST_V(12, 4, 4, 2, 2, 1, 2)
#endif
            default: return -1;
        }
    } else if constexpr (RC >= 3) {
        switch (variant) {
            LM_ST_V(19, 2, 2, 2, 2, 1, 1)
#ifdef LM_STENCIL_SHAPES
            LM_ST_X(6, 2, 2, 4, 2, 1)
#endif
            default: return -1;
        }
    } else {
        switch (variant) {
            LM_ST_V(2, 4, 2, 2, 2, 1, 1)
#ifdef LM_STENCIL_SHAPES
            LM_ST_X(13, 4, 2, 1, 2, 1) LM_ST_X(14, 4, 2, 2, 1, 1) LM_ST_X(15, 4, 2, 1, 3, 1) LM_ST_X(18, 4, 2, 1, 1, 1)
#endif
#ifdef LM_STENCIL_EXPLORE
            LM_ST_V(0, 4, 2, 2, 4, 1, 0) LM_ST_V(1, 2, 2, 2, 4, 2, 0) LM_ST_V(3, 4, 2, 2, 4, 1, 1)
            LM_ST_V(4, 2, 2, 2, 2, 2, 1) LM_ST_V(5, 2, 4, 2, 2, 1, 1) LM_ST_V(6, 2, 2, 4, 2, 1, 1)
            LM_ST_V(10, 2, 2, 4, 2, 1, 2) LM_ST_V(11, 4, 2, 2, 2, 1, 2)
#endif
            default: return -1;
        }
    }
}


// observables kernel shape per RC: T1 x T2 cells per thread, W1 x W2 warps per CTA
// (WIDE = more than 6 forward entries per row: one cell per thread keeps the partial sums in registers)
#ifndef LM_OBS_WIDE_W1
#define LM_OBS_WIDE_W1 4
#define LM_OBS_WIDE_W2 4
#endif
template <int RC, bool WIDE> struct ObsShape;
template <> struct ObsShape<1, false> { static constexpr int T1 = 2, T2 = 2, W1 = 4, W2 = 2; };   // 8 x 4 cells, 256 threads
template <> struct ObsShape<1, true>  { static constexpr int T1 = 2, T2 = 2, W1 = 4, W2 = 2; };
template <> struct ObsShape<2, false> { static constexpr int T1 = 1, T2 = 2, W1 = 4, W2 = 2; };   // 4 x 4 cells, 256 threads
template <> struct ObsShape<2, true>  { static constexpr int T1 = 1, T2 = 1, W1 = 4, W2 = 2; };   // 4 x 2 cells, 256 threads
template <> struct ObsShape<3, false> { static constexpr int T1 = 1, T2 = 2, W1 = 4, W2 = 2; };   // 3 rows x (1 + 2 NF <= 9) sums per cell
template <> struct ObsShape<3, true>  { static constexpr int T1 = 1, T2 = 1, W1 = 4, W2 = 2; };   // (512-thread CTAs measured slower on kagome NN + NNN: 3.37 -> 4.09 ms)
template <> struct ObsShape<4, false> { static constexpr int T1 = 1, T2 = 2, W1 = 4, W2 = 2; };   // 4 rows x (1 + 2 NF <= 7) sums per cell
template <> struct ObsShape<4, true>  { static constexpr int T1 = 1, T2 = 1, W1 = LM_OBS_WIDE_W1, W2 = LM_OBS_WIDE_W2; };   // one cell per thread: 4 x 4-cell patches (512 threads) keep the staged / own row ratio at 1.9 (Kane-Mele 2.34 -> 2.10 ms)
// WIDE: two cells per thread would exceed the 64 partial sums a thread folds (k_observe_stencil)
template <int RC> constexpr bool obs_wide(int nf) { return RC == 3 ? nf > 4 : (RC == 4 ? nf > 3 : nf > 6); }

template <typename T, int RC, typename MK>
static int launch_obs_t(const StencilObsArgs& a, const CUtensorMap& tmx, unsigned grid, cudaStream_t s) {
    using S = ObsShape<RC, obs_wide<RC>(st_nfwd<RC>(MK::mask))>;
    constexpr size_t smem = st_obs_smem<T, RC, S::T1, S::T2, S::W1, S::W2>();
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(k_observe_stencil<T, RC, MK, S::T1, S::T2, S::W1, S::W2>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -2;
        configured = true;
    }
    k_observe_stencil<T, RC, MK, S::T1, S::T2, S::W1, S::W2><<<grid, 32 * S::W1 * S::W2, smem, s>>>(a, tmx);
    return 0;
}
template <int RC, typename MK>
static int launch_obs(bool c64, const StencilObsArgs& a, const CUtensorMap& tmx, unsigned grid, cudaStream_t s) {
    if (!c64) return launch_obs_t<double, RC, MK>(a, tmx, grid, s);
#ifndef LM_STENCIL_NOC64
    return launch_obs_t<float, RC, MK>(a, tmx, grid, s);
#else
    return -1;
#endif
}
template <int RC, typename MK>
static void obs_shape(int* P1, int* P2, int* nf) {
    using S = ObsShape<RC, obs_wide<RC>(st_nfwd<RC>(MK::mask))>;
    *P1 = S::W1 * S::T1; *P2 = S::W2 * S::T2; *nf = st_nfwd<RC>(MK::mask);
}

}  // namespace lm
