// Register-tiled lattice-stencil kernel for the fused propagator term  y = alpha H x + gamma x (+ ...).
//
// Why: for W ~ 10 stencils (Haldane, QWZ) the ELL gather kernels are bound by the L1 data pipe -
// every (element, neighbour) pair costs its own 16-byte L1 read (profiles/r1_kernel_generations.md).
// A translation-invariant lattice (rows ordered cell-major, the reference's site order,
// src/lattices/bravais/lattice.jl:101-111: last lattice axis fastest, basis / orbital index
// innermost) lets ONE THREAD own a T1 x T2 block of unit cells of one column: every Psi element of
// the haloed block is loaded ONCE into a register and feeds all out rows it couples to, with
// compile-time register indices.  Loads per output drop from W to (T1+2)(T2+2)/(T1 T2) ~ 2.5 - 3.5.
//
// The sparsity pattern is a compile-time bit mask over the 9 cell offsets |d1|, |d2| <= 1:
//   bit (o * RC*RC + a * RC + b), o = (d1+1)*3 + (d2+1)  <=>  out row a of cell c couples to
//   in row b of cell c + (d1, d2)        (RC = rows per unit cell = basis sites x orbitals).
// Values stay per-row data (Peierls phases differ bond by bond): svals[row][slot], slots ordered
// by (o, b); entries absent on a given row (open boundaries) hold 0 and the load wraps around.
#pragma once
#include "common.cuh"
#include <utility>

namespace lm {

typedef unsigned long long st_mask_t;

template <int RC> __host__ __device__ constexpr bool st_bit(st_mask_t m, int o, int a, int b) {
    return ((m >> (o * RC * RC + a * RC + b)) & 1ull) != 0;
}
template <int RC> __host__ __device__ constexpr int st_slot(st_mask_t m, int o, int a, int b) {
    int s = 0;
    for (int oo = 0; oo < 9; ++oo)
        for (int bb = 0; bb < RC; ++bb) {
            if (oo == o && bb == b) return s;
            if (st_bit<RC>(m, oo, a, bb)) ++s;
        }
    return s;
}
template <int RC> __host__ __device__ constexpr int st_width(st_mask_t m) {
    int w = 1;
    for (int a = 0; a < RC; ++a) {
        int s = 0;
        for (int o = 0; o < 9; ++o) for (int b = 0; b < RC; ++b) if (st_bit<RC>(m, o, a, b)) ++s;
        if (s > w) w = s;
    }
    return w;
}
// is in row b of the haloed-tile cell (u1, u2), u in [0, T + 2), read by an out row of the tile?
template <int RC, int T1, int T2> __host__ __device__ constexpr bool st_needed(st_mask_t m, int u1, int u2, int b) {
    for (int o = 0; o < 9; ++o)
        for (int a = 0; a < RC; ++a)
            if (st_bit<RC>(m, o, a, b)) {
                const int v1 = u1 - 1 - (o / 3 - 1), v2 = u2 - 1 - (o % 3 - 1);
                if (v1 >= 0 && v1 < T1 && v2 >= 0 && v2 < T2) return true;
            }
    return false;
}

template <typename F, int... I>
__device__ __forceinline__ void st_for_impl(F&& f, std::integer_sequence<int, I...>) {
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N, typename F> __device__ __forceinline__ void st_for(F&& f) {
    st_for_impl(f, std::make_integer_sequence<int, N>{});
}

struct StencilArgs {
    const void* svals;              // [N][SW] complex, stencil-slot order
    int n1, n2;                     // unit cells along the slow / fast lattice axis
    int np2;                        // CTA patches along the fast axis
    long long ld;
    const void* x; void* y; const void* z; const void* u;
    double alpha[2], g[2], beta[2], delta[2];   // y = alpha (H x + g x) + beta z + delta u
    unsigned cps, nchunks;
};

__device__ __forceinline__ void pscale(double2& r, const double2 s, const double2 v) { pzero(r); pfma(r, s, v); }
__device__ __forceinline__ void pscale(float4& r, const float2 s, const float4 v) { pzero(r); pfma(r, s, v); }

__host__ __device__ constexpr int st_min_blocks(int acc_regs, int threads) {
    // accumulators + ~64 working registers per thread against the 64K-register file
    const int per_thread = acc_regs + 64;
    const int fit = 65536 / (per_thread * threads);
    return fit < 1 ? 1 : (fit > 4 ? 4 : fit);
}
__device__ __forceinline__ int st_wrap(int c, int n) { c %= n; return c < 0 ? c + n : c; }

// value-slot stride: complex64 rows are padded to an even slot count (16-byte aligned rows for TMA)
template <typename T, int RC, st_mask_t MASK> __host__ __device__ constexpr int st_stride() {
    return sizeof(T) == 4 ? ((st_width<RC>(MASK) + 1) & ~1) : st_width<RC>(MASK);
}

// The register-tile body shared by both kernels.  load_x(U1, U2, B, j) returns lane element j of
// in row B of haloed-tile cell (U1, U2); load_h(V1, V2, A, S) the value in slot S of out row A of
// tile cell (V1, V2).  All indices are integral_constants: every register index is compile-time.
template <typename T, int RC, st_mask_t MASK, int T1, int T2, int CPT, bool SELF, typename LX, typename LH>
__device__ __forceinline__ void st_tile(typename pack<T>::E (&acc)[T1][T2][RC][CPT], const typename cx2<T>::type g,
                                        LX&& load_x, LH&& load_h) {
    using T2c = typename cx2<T>::type;
    using E = typename pack<T>::E;
    st_for<T1 + 2>([&](auto U1) {
        st_for<T2 + 2>([&](auto U2) {
            st_for<RC>([&](auto B) {
                constexpr int u1 = decltype(U1)::value, u2 = decltype(U2)::value, b = decltype(B)::value;
                constexpr bool own = u1 >= 1 && u1 <= T1 && u2 >= 1 && u2 <= T2;
                constexpr bool self = own && SELF;
                if constexpr (st_needed<RC, T1, T2>(MASK, u1, u2, b) || self) {
                    E xv[CPT];
#pragma unroll
                    for (int j = 0; j < CPT; ++j) xv[j] = load_x(U1, U2, B, j);
                    st_for<9>([&](auto O) {
                        st_for<RC>([&](auto A) {
                            constexpr int o = decltype(O)::value, aa = decltype(A)::value;
                            constexpr int v1 = u1 - 1 - (o / 3 - 1), v2 = u2 - 1 - (o % 3 - 1);
                            if constexpr (st_bit<RC>(MASK, o, aa, b) && v1 >= 0 && v1 < T1 && v2 >= 0 && v2 < T2) {
                                constexpr int slot = st_slot<RC>(MASK, o, aa, b);
                                const T2c hv = load_h(std::integral_constant<int, v1>{}, std::integral_constant<int, v2>{},
                                                      A, std::integral_constant<int, slot>{});
#pragma unroll
                                for (int j = 0; j < CPT; ++j) pfma(acc[v1][v2][aa][j], hv, xv[j]);
                            }
                        });
                    });
                    if constexpr (self) {
#pragma unroll
                        for (int j = 0; j < CPT; ++j) pfma(acc[u1 - 1][u2 - 1][b][j], g, xv[j]);
                    }
                }
            });
        });
    });
}

// ------------------------------------------------------------------------------------------
// k_apply_stencil: direct variant - the haloed block is read with ld.global.nc (L1 shares the halo
// rows between the W1 x W2 register tiles of a CTA).  Latency-bound in practice: the loads in
// flight are limited by the registers the accumulators leave over (profiles/).
// ------------------------------------------------------------------------------------------
template <typename T, int RC, st_mask_t MASK, int T1, int T2, int W1, int W2, int CPT, int MODE>
__global__ void __launch_bounds__(32 * W1 * W2, st_min_blocks(T1 * T2 * RC * CPT * 4, 32 * W1 * W2))
k_apply_stencil(const StencilArgs a) {
    using T2c = typename cx2<T>::type;
    using E = typename pack<T>::E;
    constexpr int EC = pack<T>::EC;
    constexpr int SWP = st_stride<T, RC, MASK>();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned patch = blockIdx.x / a.cps;
    const unsigned chunk = blockIdx.y * a.cps + (blockIdx.x - patch * a.cps);
    if (chunk >= a.nchunks) return;
    const int pj1 = (int)(patch / (unsigned)a.np2), pj2 = (int)(patch - (unsigned)pj1 * (unsigned)a.np2);
    const int o1 = (pj1 * W1 + warp / W2) * T1, o2 = (pj2 * W2 + warp % W2) * T2;
    if (o1 >= a.n1 || o2 >= a.n2) return;
    const long long lde = a.ld / EC;
    const E* __restrict__ x = (const E*)a.x;
    const T2c* __restrict__ sv = (const T2c*)a.svals;
    long long cidx[CPT];
    bool ok[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        const long long c = (long long)chunk * (32 * CPT) + lane + 32 * j;
        ok[j] = c < lde;
        cidx[j] = ok[j] ? c : (lde - 1);
    }
    // wrapped cell coordinates of the haloed tile (periodic images; open-boundary entries are 0)
    int r1[T1 + 2], r2[T2 + 2];
#pragma unroll
    for (int u = 0; u < T1 + 2; ++u) r1[u] = st_wrap(o1 + u - 1, a.n1) * a.n2;
#pragma unroll
    for (int u = 0; u < T2 + 2; ++u) r2[u] = st_wrap(o2 + u - 1, a.n2);
    E acc[T1][T2][RC][CPT];
#pragma unroll
    for (int v1 = 0; v1 < T1; ++v1)
#pragma unroll
        for (int v2 = 0; v2 < T2; ++v2)
#pragma unroll
            for (int aa = 0; aa < RC; ++aa)
#pragma unroll
                for (int j = 0; j < CPT; ++j) pzero(acc[v1][v2][aa][j]);
    const T2c g = cmake<T2c>(a.g[0], a.g[1]);

    st_tile<T, RC, MASK, T1, T2, CPT, (MODE == 2 || MODE == 3)>(acc, g,
        [&](auto U1, auto U2, auto B, int j) {
            const long long row = (long long)(r1[decltype(U1)::value] + r2[decltype(U2)::value]) * RC + decltype(B)::value;
            return ld_ro(x + row * lde + cidx[j]);
        },
        [&](auto V1, auto V2, auto A, auto S) {
            return sv[((long long)(r1[decltype(V1)::value + 1] + r2[decltype(V2)::value + 1]) * RC + decltype(A)::value) * SWP + decltype(S)::value];
        });

    const T2c alpha = cmake<T2c>(a.alpha[0], a.alpha[1]);
    const T2c beta  = cmake<T2c>(a.beta[0],  a.beta[1]);
    const T2c delta = cmake<T2c>(a.delta[0], a.delta[1]);
    E* y = (E*)a.y;
    const E* z = (const E*)a.z;
    const E* u = (const E*)a.u;
#pragma unroll
    for (int v1 = 0; v1 < T1; ++v1)
#pragma unroll
        for (int v2 = 0; v2 < T2; ++v2) {
            if (o1 + v1 >= a.n1 || o2 + v2 >= a.n2) continue;      // ragged last tile
#pragma unroll
            for (int aa = 0; aa < RC; ++aa) {
                const long long row = (long long)(r1[v1 + 1] + r2[v2 + 1]) * RC + aa;
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    const long long e = row * lde + cidx[j];
                    E res;
                    pscale(res, alpha, acc[v1][v2][aa][j]);
                    if (MODE == 1) pfma(res, beta, ld_stream(z + e));
                    if (MODE == 2) {
                        if (z) pfma(res, beta, ld_stream(z + e));
                        if (u) pfma(res, delta, u[e]);
                    }
                    if (ok[j]) st_stream(y + e, res);
                }
            }
        }
}

// ------------------------------------------------------------------------------------------
// k_apply_stencil_tma: staged variant.  A CTA owns a patch of (W1 T1) x (W2 T2) cells and one chunk
// of 32 CPT lane elements: the haloed Psi rows of the chunk (one cp.async.bulk per row) and the
// value rows of the patch (one bulk copy per cell line) are brought into shared memory by the TMA
// engine - every byte of the CTA is in flight at once, no registers - and each warp then runs its
// T1 x T2 register tile out of shared memory (128-bit LDS at compile-time offsets, the value
// loads are broadcasts).  Several CTAs per SM overlap one CTA's copies with another's FMAs.
// ------------------------------------------------------------------------------------------
template <typename T, int RC, st_mask_t MASK, int T1, int T2, int W1, int W2, int CPT>
__host__ __device__ constexpr size_t st_tma_smem() {
    return (size_t)(W1 * T1 + 2) * (W2 * T2 + 2) * RC * 32 * CPT * 16
         + (size_t)(W1 * T1) * (W2 * T2) * RC * st_stride<T, RC, MASK>() * (2 * sizeof(T));
}
template <typename T, int RC, st_mask_t MASK, int T1, int T2, int W1, int W2, int CPT>
__host__ __device__ constexpr int st_tma_blocks() {
    const int by_smem = (int)((227 * 1024) / (st_tma_smem<T, RC, MASK, T1, T2, W1, W2, CPT>() + 1024 + 64));
    const int by_regs = st_min_blocks(T1 * T2 * RC * CPT * 4, 32 * W1 * W2);
    const int m = by_smem < by_regs ? by_smem : by_regs;
    return m < 1 ? 1 : m;
}

template <typename T, int RC, st_mask_t MASK, int T1, int T2, int W1, int W2, int CPT, int MODE>
__global__ void __launch_bounds__(32 * W1 * W2, st_tma_blocks<T, RC, MASK, T1, T2, W1, W2, CPT>())
k_apply_stencil_tma(const StencilArgs a) {
    using T2c = typename cx2<T>::type;
    using E = typename pack<T>::E;
    constexpr int EC = pack<T>::EC;
    constexpr int NT = 32 * W1 * W2, P1 = W1 * T1, P2 = W2 * T2;
    constexpr int HR = (P1 + 2) * (P2 + 2) * RC;            // haloed rows of the patch
    constexpr int CE = 32 * CPT;                            // lane elements per staged row
    constexpr int SWP = st_stride<T, RC, MASK>();
    extern __shared__ __align__(128) unsigned char lm_smem[];
    E* sx = reinterpret_cast<E*>(lm_smem);                                   // [HR][CE]
    T2c* sh = reinterpret_cast<T2c*>(lm_smem + (size_t)HR * CE * sizeof(E));  // [P1][P2 * RC * SWP]
    __shared__ __align__(8) unsigned long long bar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned patch = blockIdx.x / a.cps;
    const unsigned chunk = blockIdx.y * a.cps + (blockIdx.x - patch * a.cps);
    if (chunk >= a.nchunks) return;
    const int pj1 = (int)(patch / (unsigned)a.np2), pj2 = (int)(patch - (unsigned)pj1 * (unsigned)a.np2);
    const int o1 = pj1 * P1, o2 = pj2 * P2;                 // patch origin (always inside the lattice)
    const long long lde = a.ld / EC;
    const long long c0 = (long long)chunk * CE;
    const int cw = (int)((lde - c0) < CE ? (lde - c0) : CE);
    const int vl1 = (a.n1 - o1) < P1 ? (a.n1 - o1) : P1;    // own cells inside the lattice
    const int vl2 = (a.n2 - o2) < P2 ? (a.n2 - o2) : P2;
    const unsigned hline = ((unsigned)(vl2 * RC * SWP * (int)sizeof(T2c)) + 15u) & ~15u;
    const E* __restrict__ x = (const E*)a.x;
    const T2c* __restrict__ sv = (const T2c*)a.svals;
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_arrive_expect_tx(&bar, (unsigned)(HR * cw * (int)sizeof(E)) + (unsigned)vl1 * hline);
    }
    __syncthreads();
    for (int r = tid; r < HR; r += NT) {
        const int u1 = r / ((P2 + 2) * RC), rem = r - u1 * ((P2 + 2) * RC), u2 = rem / RC, b = rem - u2 * RC;
        const long long row = ((long long)st_wrap(o1 + u1 - 1, a.n1) * a.n2 + st_wrap(o2 + u2 - 1, a.n2)) * RC + b;
        tma_bulk_g2s(sx + r * CE, x + row * lde + c0, (unsigned)(cw * (int)sizeof(E)), &bar);
    }
    for (int l = NT - 1 - tid; l < vl1; l += NT)
        tma_bulk_g2s(sh + l * (P2 * RC * SWP), sv + ((long long)(o1 + l) * a.n2 + o2) * (RC * SWP), hline, &bar);

    const int w1 = warp / W2, w2 = warp % W2;
    const int q1 = o1 + w1 * T1, q2 = o2 + w2 * T2;         // first cell of this warp's register tile
    const E* xb = sx + ((w1 * T1) * (P2 + 2) + w2 * T2) * RC * CE + lane;
    const T2c* hb = sh + ((w1 * T1) * P2 + w2 * T2) * RC * SWP;
    E acc[T1][T2][RC][CPT];
#pragma unroll
    for (int v1 = 0; v1 < T1; ++v1)
#pragma unroll
        for (int v2 = 0; v2 < T2; ++v2)
#pragma unroll
            for (int aa = 0; aa < RC; ++aa)
#pragma unroll
                for (int j = 0; j < CPT; ++j) pzero(acc[v1][v2][aa][j]);
    const T2c g = cmake<T2c>(a.g[0], a.g[1]);

    mbar_wait(&bar, 0);
    if (q1 >= a.n1 || q2 >= a.n2) return;

    st_tile<T, RC, MASK, T1, T2, CPT, (MODE == 2 || MODE == 3)>(acc, g,
        [&](auto U1, auto U2, auto B, int j) {
            return xb[((decltype(U1)::value * (P2 + 2) + decltype(U2)::value) * RC + decltype(B)::value) * CE + 32 * j];
        },
        [&](auto V1, auto V2, auto A, auto S) {
            return hb[((decltype(V1)::value * P2 + decltype(V2)::value) * RC + decltype(A)::value) * SWP + decltype(S)::value];
        });

    const T2c alpha = cmake<T2c>(a.alpha[0], a.alpha[1]);
    const T2c beta  = cmake<T2c>(a.beta[0],  a.beta[1]);
    const T2c delta = cmake<T2c>(a.delta[0], a.delta[1]);
    E* y = (E*)a.y;
    const E* z = (const E*)a.z;
    const E* u = (const E*)a.u;
#pragma unroll
    for (int v1 = 0; v1 < T1; ++v1)
#pragma unroll
        for (int v2 = 0; v2 < T2; ++v2) {
            if (q1 + v1 >= a.n1 || q2 + v2 >= a.n2) continue;      // ragged last tile
#pragma unroll
            for (int aa = 0; aa < RC; ++aa) {
                const long long row = ((long long)(q1 + v1) * a.n2 + (q2 + v2)) * RC + aa;
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    const long long cj = c0 + lane + 32 * j;
                    if (cj >= lde) continue;
                    const long long e = row * lde + cj;
                    E res;
                    pscale(res, alpha, acc[v1][v2][aa][j]);
                    if (MODE == 1) pfma(res, beta, ld_stream(z + e));
                    if (MODE == 2) {
                        if (z) pfma(res, beta, ld_stream(z + e));
                        if (u) pfma(res, delta, u[e]);
                    }
                    st_stream(y + e, res);
                }
            }
        }
}

// ---- compiled patterns (stencil.cu registry; one translation unit per pattern) ----
#define LM_ST_MASK0 0xbaull
#define LM_ST_MASK1 0x1ffull
#define LM_ST_MASK2 0x404f2020ull
#define LM_ST_MASK3 0xf0fff0f0ull
#define LM_ST_MASK4 0xd9dfb9b0ull

// ---- host-visible registry (stencil.cu) ----
struct StencilDesc { int rc; st_mask_t mask; int sw; const char* name; };
int stencil_count();
const StencilDesc& stencil_desc(int id);
int stencil_find(int rc, st_mask_t mask);                  // smallest compiled superset, -1 if none
int stencil_stride(int id, bool c64);                      // value-slot stride (complex64 rows are padded to even)
int stencil_num_variants();
// patch size in cells, lane elements per thread and kernel family (staged = TMA) of a variant
void stencil_variant_shape(int variant, int* P1, int* P2, int* cpt, int* staged);
// launches; returns 0 on success, -1 if (id, variant, mode) is not compiled, -2 on a CUDA error
int stencil_launch(int id, int variant, bool c64, int mode, const StencilArgs& a, dim3 grid, cudaStream_t s);

}  // namespace lm
