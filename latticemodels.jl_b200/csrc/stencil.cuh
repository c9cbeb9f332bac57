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
//   in row b of cell c + (d1, d2)        (RC = rows per unit cell = basis sites x orbitals, <= 4).
// A compiled pattern is a TYPE MK (struct StPat<id> at the end of this file) that carries the mask
// and the value classes as static constexpr members: a class-type template argument cannot be
// used for a __global__ function (nvcc 12.9), a type can.
// Values stay per-row data (Peierls phases differ bond by bond): svals[row][slot], slots ordered
// by (o, b); entries absent on a given row (open boundaries) hold 0 and the load wraps around.
#pragma once
#include "common.cuh"

namespace lm {

// 256 pattern bits: RC <= 4 rows per cell at the 9 offsets (144 bits)
struct st_mask_t { unsigned long long w[4]; };
__host__ __device__ constexpr bool st_get(const st_mask_t& m, int i) { return ((m.w[i >> 6] >> (i & 63)) & 1ull) != 0; }
__host__ __device__ constexpr bool st_covers(const st_mask_t& big, const st_mask_t& small) {
    for (int k = 0; k < 4; ++k) if (small.w[k] & ~big.w[k]) return false;
    return true;
}
#ifndef __CUDACC_RTC__
inline void st_set(st_mask_t& m, int i) { m.w[i >> 6] |= 1ull << (i & 63); }
#endif

template <int RC> __host__ __device__ constexpr bool st_bit(const st_mask_t& m, int o, int a, int b) {
    return st_get(m, o * RC * RC + a * RC + b);
}
template <int RC> __host__ __device__ constexpr int st_slot(const st_mask_t& m, int o, int a, int b) {
    int s = 0;
    for (int oo = 0; oo < 9; ++oo)
        for (int bb = 0; bb < RC; ++bb) {
            if (oo == o && bb == b) return s;
            if (st_bit<RC>(m, oo, a, bb)) ++s;
        }
    return s;
}
template <int RC> __host__ __device__ constexpr int st_width(const st_mask_t& m) {
    int w = 1;
    for (int a = 0; a < RC; ++a) {
        int s = 0;
        for (int o = 0; o < 9; ++o) for (int b = 0; b < RC; ++b) if (st_bit<RC>(m, o, a, b)) ++s;
        if (s > w) w = s;
    }
    return w;
}
// is in row b of the haloed-tile cell (u1, u2), u in [0, T + 2), read by an out row of the tile?
template <int RC, int T1, int T2> __host__ __device__ constexpr bool st_needed(const st_mask_t& m, int u1, int u2, int b) {
    for (int o = 0; o < 9; ++o)
        for (int a = 0; a < RC; ++a)
            if (st_bit<RC>(m, o, a, b)) {
                const int v1 = u1 - 1 - (o / 3 - 1), v2 = u2 - 1 - (o % 3 - 1);
                if (v1 >= 0 && v1 < T1 && v2 >= 0 && v2 < T2) return true;
            }
    return false;
}

// compile-time loop: f(st_ic<0>{}), ..., f(st_ic<N - 1>{})  (own integer sequence: the header also compiles under NVRTC, which has no <utility>)
template <int V> struct st_ic { static constexpr int value = V; };
template <int... I> struct st_iseq {};
template <int N, int... I> struct st_mkseq : st_mkseq<N - 1, N - 1, I...> {};
template <int... I> struct st_mkseq<0, I...> { using type = st_iseq<I...>; };
template <typename F, int... I>
__device__ __forceinline__ void st_for_impl(F&& f, st_iseq<I...>) {
    (f(st_ic<I>{}), ...);
}
template <int N, typename F> __device__ __forceinline__ void st_for(F&& f) {
    st_for_impl(f, typename st_mkseq<N>::type{});
}

struct StencilArgs {
    const void* svals;              // [N][SW] complex, stencil-slot order
    const void* sreal;              // [N][SWR] real scalars, same slot order: the values in the pattern's real / imaginary class
                                    //    (Re of a real-class entry, Im of an imaginary-class entry; st_rstride)
    const int* ri_flag;             // device flag (null = never): non-zero iff the CURRENT values are in that class (k_gather_real
                                    //    after every value change) - read by the kernel, so a replayed step graph follows the values
    int n1, n2;                     // unit cells along the slow / fast lattice axis
    int np2;                        // CTA patches along the fast axis
    long long ld;                   // row stride of x / y / z / u (complex elements)
    int pdl;                        // 1: launched with programmatic stream serialization (k_apply_stencil_tma only):
                                    //    let the next launch of the chain start its CTAs while this grid drains,
                                    //    and wait for the previous grid before touching global memory
    int herm;                       // 1: H is Hermitian (checked on the device): in-tile bonds share one value load (st_tile_herm)
    int tmap;                       // 1: interior patches stage their haloed block with ONE 3-D tensor-map copy (k_apply_stencil_tma)
    unsigned pf;                    // > 0: every CTA also prefetches into L2 the box of the CTA `pf` positions later in dispatch order
    const void* x; void* y; const void* z; const void* u;
    double alpha[2], g[2], beta[2], delta[2];   // y = alpha (H x + g x) + beta z + delta u
    unsigned cps, nchunks;
    unsigned ngroups, cpg, npatch;  // streaming kernel: column groups per patch, chunks per group, patches
};

__device__ __forceinline__ void pscale(double2& r, const double2 s, const double2 v) { pzero(r); pfma(r, s, v); }
__device__ __forceinline__ void pscale(float4& r, const float2 s, const float4 v) { pzero(r); pfma(r, s, v); }

__host__ __device__ constexpr int st_min_blocks(int acc_regs, int threads) {
    // accumulators + ~64 working registers per thread against the 64K-register file
    const int per_thread = acc_regs + 64;
    const int fit = 65536 / (per_thread * threads);
    return fit < 1 ? 1 : (fit > 8 ? 8 : fit);
}
__device__ __forceinline__ int st_wrap(int c, int n) { c %= n; return c < 0 ? c + n : c; }

// value-slot stride: complex64 rows are padded to an even slot count (16-byte aligned rows for TMA)
template <typename T, int RC, typename MK> __host__ __device__ constexpr int st_stride() {
    return sizeof(T) == 4 ? ((st_width<RC>(MK::mask) + 1) & ~1) : st_width<RC>(MK::mask);
}

// scalars per row of the real / imaginary class copy: rows are 16-byte multiples (bulk copies of whole cell lines)
template <typename T, int RC, typename MK> __host__ __device__ constexpr int st_rstride() {
    return sizeof(T) == 4 ? ((st_width<RC>(MK::mask) + 3) & ~3) : ((st_width<RC>(MK::mask) + 1) & ~1);
}
// one entry: acc += h x (CONJ: conj(h) x).  CLS 0: h complex; 1: h = s purely real; 2: h = i s purely imaginary
template <int CLS, bool CONJ, typename HV, typename E>
__device__ __forceinline__ void st_fma(E& acc, const HV hv, const E x) {
    if constexpr (CLS == 0) { if constexpr (CONJ) pfma_conj(acc, hv, x); else pfma(acc, hv, x); }
    else if constexpr (CLS == 1) pfma_re(acc, hv, x);
    else pfma_im(acc, CONJ ? -hv : hv, x);
}
// value class of entry (o, a, b) in a kernel that reads class scalars (VC = 1) or complex values (VC = 0)
template <int RC, typename MK, int VC> __host__ __device__ constexpr int st_cls(int o, int a, int b) {
    return VC ? (st_bit<RC>(MK::imag, o, a, b) ? 2 : 1) : 0;
}

// The register-tile body shared by both kernels.  load_x(U1, U2, B, j) returns lane element j of
// in row B of haloed-tile cell (U1, U2); load_h(V1, V2, A, S) the value in slot S of out row A of
// tile cell (V1, V2).  All indices are integral_constants: every register index is compile-time.
template <typename T, int RC, typename MK, int T1, int T2, int CPT, bool SELF, int VC, typename LX, typename LH>
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
                if constexpr (st_needed<RC, T1, T2>(MK::mask, u1, u2, b) || self) {
                    E xv[CPT];
#pragma unroll
                    for (int j = 0; j < CPT; ++j) xv[j] = load_x(U1, U2, B, j);
                    st_for<9>([&](auto O) {
                        st_for<RC>([&](auto A) {
                            constexpr int o = decltype(O)::value, aa = decltype(A)::value;
                            constexpr int v1 = u1 - 1 - (o / 3 - 1), v2 = u2 - 1 - (o % 3 - 1);
                            if constexpr (st_bit<RC>(MK::mask, o, aa, b) && v1 >= 0 && v1 < T1 && v2 >= 0 && v2 < T2) {
                                constexpr int slot = st_slot<RC>(MK::mask, o, aa, b);
                                const auto hv = load_h(st_ic<v1>{}, st_ic<v2>{},
                                                       A, st_ic<slot>{});
#pragma unroll
                                for (int j = 0; j < CPT; ++j) st_fma<st_cls<RC, MK, VC>(o, aa, b), false>(acc[v1][v2][aa][j], hv, xv[j]);
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

// Is the pattern structurally symmetric (entry (a <- b at offset d) present iff (b <- a at -d) is)?
template <int RC> __host__ __device__ constexpr bool st_symmetric(const st_mask_t& m) {
    for (int o = 0; o < 9; ++o)
        for (int a = 0; a < RC; ++a)
            for (int b = 0; b < RC; ++b)
                if (st_bit<RC>(m, o, a, b) != st_bit<RC>(m, 8 - o, b, a)) return false;
    return true;
}
// forward half of a symmetric pattern: offset (+1, *), (0, +1), or a later row of the same cell
template <int RC> __host__ __device__ constexpr bool st_fwd_bit(const st_mask_t& m, int o, int a, int b) {
    return st_bit<RC>(m, o, a, b) && (o > 4 || (o == 4 && b > a));
}

// Register-tile body for a HERMITIAN operator (H[q,p] = conj(H[p,q]); checked on the device whenever
// the values change, api.cu).  The uniform 128-bit value loads are the largest share of the kernel's
// shared-memory wavefronts (each costs two; profiles/r1_stencil_kernels.md), so a bond whose two rows
// both live in the thread's tile loads its value ONCE and feeds both directions:
//     acc[p] += h x[q] ,   acc[q] += conj(h) x[p]        (Haldane 4 x 2 tile: 116 value loads instead of 160).
// The own rows of the tile are held in registers; halo rows are loaded one at a time and feed the
// own rows they couple to, as in st_tile.  `g` is folded into the diagonal value (2 additions per
// row and thread instead of 4 FMAs per element).  Only valid for a tile that lies completely
// inside the lattice (the value rows of cells beyond the edge are not staged): the kernels fall
// back to st_tile for ragged tiles.
template <typename T, int RC, typename MK, int T1, int T2, int CPT, bool SELF, int VC, typename LX, typename LH>
__device__ __forceinline__ void st_tile_herm(typename pack<T>::E (&acc)[T1][T2][RC][CPT], const typename cx2<T>::type g,
                                             LX&& load_x, LH&& load_h) {
    using T2c = typename cx2<T>::type;
    using E = typename pack<T>::E;
    static_assert(st_symmetric<RC>(MK::mask), "st_tile_herm needs a structurally symmetric pattern");
    E xo[T1][T2][RC][CPT];
    st_for<T1>([&](auto V1) { st_for<T2>([&](auto V2) { st_for<RC>([&](auto A) {
        constexpr int v1 = decltype(V1)::value, v2 = decltype(V2)::value, aa = decltype(A)::value;
#pragma unroll
        for (int j = 0; j < CPT; ++j)
            xo[v1][v2][aa][j] = load_x(st_ic<v1 + 1>{}, st_ic<v2 + 1>{}, A, j);
    }); }); });
    // diagonal entries (+ g) and the bonds inside the tile, one value load per bond
    st_for<T1>([&](auto V1) { st_for<T2>([&](auto V2) { st_for<RC>([&](auto A) {
        constexpr int v1 = decltype(V1)::value, v2 = decltype(V2)::value, aa = decltype(A)::value;
        if constexpr (st_bit<RC>(MK::mask, 4, aa, aa)) {
            const auto h0 = load_h(V1, V2, A, st_ic<st_slot<RC>(MK::mask, 4, aa, aa)>{});
            if constexpr (VC && !SELF) {
                static_assert(!VC || !st_bit<RC>(MK::imag, 4, aa, aa), "the diagonal of a Hermitian operator is real");
#pragma unroll
                for (int j = 0; j < CPT; ++j) pfma_re(acc[v1][v2][aa][j], h0, xo[v1][v2][aa][j]);
            } else {
                T2c hv;
                if constexpr (VC) { hv.x = h0; hv.y = 0; } else hv = h0;
                if constexpr (SELF) { hv.x += g.x; hv.y += g.y; }
#pragma unroll
                for (int j = 0; j < CPT; ++j) pfma(acc[v1][v2][aa][j], hv, xo[v1][v2][aa][j]);
            }
        } else if constexpr (SELF) {
#pragma unroll
            for (int j = 0; j < CPT; ++j) pfma(acc[v1][v2][aa][j], g, xo[v1][v2][aa][j]);
        }
        st_for<5>([&](auto OO) { st_for<RC>([&](auto B) {
            constexpr int o = decltype(OO)::value + 4, b = decltype(B)::value;
            constexpr int w1 = v1 + (o / 3 - 1), w2 = v2 + (o % 3 - 1);
            if constexpr (st_fwd_bit<RC>(MK::mask, o, aa, b) && w1 >= 0 && w1 < T1 && w2 >= 0 && w2 < T2) {
                const auto hv = load_h(V1, V2, A, st_ic<st_slot<RC>(MK::mask, o, aa, b)>{});
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    st_fma<st_cls<RC, MK, VC>(o, aa, b), false>(acc[v1][v2][aa][j], hv, xo[w1][w2][b][j]);
                    st_fma<st_cls<RC, MK, VC>(o, aa, b), true>(acc[w1][w2][b][j], hv, xo[v1][v2][aa][j]);
                }
            }
        }); });
    }); }); });
    // halo rows: loaded once, feed every own row they couple to
    st_for<T1 + 2>([&](auto U1) { st_for<T2 + 2>([&](auto U2) { st_for<RC>([&](auto B) {
        constexpr int u1 = decltype(U1)::value, u2 = decltype(U2)::value, b = decltype(B)::value;
        constexpr bool own = u1 >= 1 && u1 <= T1 && u2 >= 1 && u2 <= T2;
        if constexpr (!own && st_needed<RC, T1, T2>(MK::mask, u1, u2, b)) {
            E xv[CPT];
#pragma unroll
            for (int j = 0; j < CPT; ++j) xv[j] = load_x(U1, U2, B, j);
            st_for<9>([&](auto O) { st_for<RC>([&](auto A) {
                constexpr int o = decltype(O)::value, aa = decltype(A)::value;
                constexpr int v1 = u1 - 1 - (o / 3 - 1), v2 = u2 - 1 - (o % 3 - 1);
                if constexpr (st_bit<RC>(MK::mask, o, aa, b) && v1 >= 0 && v1 < T1 && v2 >= 0 && v2 < T2) {
                    const auto hv = load_h(st_ic<v1>{}, st_ic<v2>{},
                                           A, st_ic<st_slot<RC>(MK::mask, o, aa, b)>{});
#pragma unroll
                    for (int j = 0; j < CPT; ++j) st_fma<st_cls<RC, MK, VC>(o, aa, b), false>(acc[v1][v2][aa][j], hv, xv[j]);
                }
            }); });
        }
    }); }); });
}

// ------------------------------------------------------------------------------------------
// k_apply_stencil: direct variant - the haloed block is read with ld.global.nc (L1 shares the halo
// rows between the W1 x W2 register tiles of a CTA).  Latency-bound in practice: the loads in
// flight are limited by the registers the accumulators leave over (profiles/).
// ------------------------------------------------------------------------------------------
template <typename T, int RC, typename MK, int T1, int T2, int W1, int W2, int CPT, int MODE>
__global__ void __launch_bounds__(32 * W1 * W2, st_min_blocks(T1 * T2 * RC * CPT * 4, 32 * W1 * W2))
k_apply_stencil(const StencilArgs a) {
    using T2c = typename cx2<T>::type;
    using E = typename pack<T>::E;
    constexpr int EC = pack<T>::EC;
    constexpr int SWP = st_stride<T, RC, MK>();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned patch = blockIdx.x / a.cps;
    const unsigned chunk = blockIdx.y * a.cps + (blockIdx.x - patch * a.cps);
    if (chunk >= a.nchunks) return;
    const int pj1 = (int)(patch / (unsigned)a.np2), pj2 = (int)(patch - (unsigned)pj1 * (unsigned)a.np2);
    const int o1 = (pj1 * W1 + warp / W2) * T1, o2 = (pj2 * W2 + warp % W2) * T2;
    if (o1 >= a.n1 || o2 >= a.n2) return;
    const long long lde = a.ld / EC, nce = lde;
    const E* __restrict__ x = (const E*)a.x;
    const T2c* __restrict__ sv = (const T2c*)a.svals;
    long long cidx[CPT];
    bool ok[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        const long long c = (long long)chunk * (32 * CPT) + lane + 32 * j;
        ok[j] = c < nce;
        cidx[j] = ok[j] ? c : (nce - 1);
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

    st_tile<T, RC, MK, T1, T2, CPT, (MODE == 2 || MODE == 3), 0>(acc, g,
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
template <typename T, int RC, typename MK, int T1, int T2, int W1, int W2, int CPT>
__host__ __device__ constexpr size_t st_tma_smem() {
    return (size_t)(W1 * T1 + 2) * (W2 * T2 + 2) * RC * 32 * CPT * 16
         + (size_t)(W1 * T1) * (W2 * T2) * RC * st_stride<T, RC, MK>() * (2 * sizeof(T));
}
template <typename T, int RC, typename MK, int T1, int T2, int W1, int W2, int CPT>
__host__ __device__ constexpr int st_tma_blocks() {
    const int by_smem = (int)((227 * 1024) / (st_tma_smem<T, RC, MK, T1, T2, W1, W2, CPT>() + 1024 + 64));
    const int by_regs = st_min_blocks(T1 * T2 * RC * CPT * 4, 32 * W1 * W2);
    const int m = by_smem < by_regs ? by_smem : by_regs;
    return m < 1 ? 1 : m;
}

template <typename T, int RC, typename MK, int T1, int T2, int W1, int W2, int CPT, int MODE>
__global__ void __launch_bounds__(32 * W1 * W2, st_tma_blocks<T, RC, MK, T1, T2, W1, W2, CPT>())
k_apply_stencil_tma(const StencilArgs a, const LM_GRID_CONSTANT CUtensorMap tmx) {
    using T2c = typename cx2<T>::type;
    using E = typename pack<T>::E;
    constexpr int EC = pack<T>::EC;
    constexpr int NT = 32 * W1 * W2, P1 = W1 * T1, P2 = W2 * T2;
    constexpr int HR = (P1 + 2) * (P2 + 2) * RC;            // haloed rows of the patch
    constexpr int CE = 32 * CPT;                            // lane elements per staged row
    constexpr int SWP = st_stride<T, RC, MK>();
    constexpr int SWR = st_rstride<T, RC, MK>();
    LM_SMEM_DYN(lm_smem);
    E* sx = reinterpret_cast<E*>(lm_smem);                                   // [HR][CE]
    T2c* sh = reinterpret_cast<T2c*>(lm_smem + (size_t)HR * CE * sizeof(E));  // [P1][P2 * RC * SWP]  (class scalars: [P1][P2 * RC * SWR] of T)
    LM_SMEM_STATIC __align__(8) unsigned long long bar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned patch = blockIdx.x / a.cps;
    const unsigned chunk = blockIdx.y * a.cps + (blockIdx.x - patch * a.cps);
    if (chunk >= a.nchunks) return;
    const int pj1 = (int)(patch / (unsigned)a.np2), pj2 = (int)(patch - (unsigned)pj1 * (unsigned)a.np2);
    const int o1 = pj1 * P1, o2 = pj2 * P2;                 // patch origin (always inside the lattice)
    const long long lde = a.ld / EC, nce = lde;
    const long long c0 = (long long)chunk * CE;
    const int cw = (int)((nce - c0) < CE ? (nce - c0) : CE);
    const int vl1 = (a.n1 - o1) < P1 ? (a.n1 - o1) : P1;    // own cells inside the lattice
    const int vl2 = (a.n2 - o2) < P2 ? (a.n2 - o2) : P2;
    const E* __restrict__ x = (const E*)a.x;
    const T2c* __restrict__ sv = (const T2c*)a.svals;
    // a patch whose haloed block lies inside the lattice is one dense box of the [n1][n2 RC][columns]
    // view of x: ONE tensor-map copy (SASS UTMALDG; columns beyond the block are zero-filled and
    // counted) instead of HR row copies.  Patches on the rim keep the row copies (periodic images).
    const bool boxed = a.tmap && o1 >= 1 && o1 + P1 + 1 <= a.n1 && o2 >= 1 && o2 + P2 + 1 <= a.n2;
    if (tid == 0) mbar_init(&bar, 1);
    // a chain of factors (x <- y of the previous launch): with programmatic dependent launch the
    // CTAs of this grid are scheduled while the previous grid drains; nothing of global memory is
    // touched before the previous grid has completed and flushed
    if (a.pdl) { pdl_launch_dependents(); pdl_wait(); }
    // real / imaginary value class (grid-uniform): the values are staged as scalars, half the bytes
    const bool ri = a.ri_flag != nullptr && *reinterpret_cast<const volatile int*>(a.ri_flag) != 0;
    const unsigned hline = ri ? (((unsigned)(vl2 * RC * SWR * (int)sizeof(T)) + 15u) & ~15u)
                              : (((unsigned)(vl2 * RC * SWP * (int)sizeof(T2c)) + 15u) & ~15u);
    if (tid == 0) mbar_arrive_expect_tx(&bar, (unsigned)(HR * (boxed ? CE : cw) * (int)sizeof(E)) + (unsigned)vl1 * hline);
    __syncthreads();
    if (boxed) {
        if (tid == 0) tma_tensor3d_g2s(sx, &tmx, (int)(c0 * (long long)(sizeof(E) / 8)), (o2 - 1) * RC, o1 - 1, &bar);
    } else {
        for (int r = tid; r < HR; r += NT) {
            const int u1 = r / ((P2 + 2) * RC), rem = r - u1 * ((P2 + 2) * RC), u2 = rem / RC, b = rem - u2 * RC;
            const long long row = ((long long)st_wrap(o1 + u1 - 1, a.n1) * a.n2 + st_wrap(o2 + u2 - 1, a.n2)) * RC + b;
            tma_bulk_g2s(sx + r * CE, x + row * lde + c0, (unsigned)(cw * (int)sizeof(E)), &bar);
        }
    }
    if (ri) {
        const T* __restrict__ sr = (const T*)a.sreal;
        for (int l = NT - 1 - tid; l < vl1; l += NT)
            tma_bulk_g2s(reinterpret_cast<T*>(sh) + l * (P2 * RC * SWR), sr + ((long long)(o1 + l) * a.n2 + o2) * (RC * SWR), hline, &bar);
    } else {
        for (int l = NT - 1 - tid; l < vl1; l += NT)
            tma_bulk_g2s(sh + l * (P2 * RC * SWP), sv + ((long long)(o1 + l) * a.n2 + o2) * (RC * SWP), hline, &bar);
    }
    // The wait for the staged block is the largest stall of this kernel (42 % of the warp samples, profiles/r2/).
    // Opt-in (a.pf > 0, LM_STENCIL_PF): every CTA also asks the TMA engine to pull into L2 the box of the CTA that
    // will be dispatched `pf` positions later, whose own copy is then an L2 hit.  A hint only: nothing waits on
    // it.  Measured neutral (the wait shrinks, the kernel time does not): off by default.
    if (a.pf && a.tmap && tid == 32 % NT) {
        const unsigned long long lin = (unsigned long long)blockIdx.y * gridDim.x + blockIdx.x + a.pf;
        if (lin < (unsigned long long)gridDim.x * gridDim.y) {
            const unsigned bx = (unsigned)(lin % gridDim.x), by = (unsigned)(lin / gridDim.x);
            const unsigned patch2 = bx / a.cps, chunk2 = by * a.cps + (bx - patch2 * a.cps);
            const int p1 = (int)(patch2 / (unsigned)a.np2) * P1, p2 = (int)(patch2 % (unsigned)a.np2) * P2;
            if (chunk2 < a.nchunks && p1 >= 1 && p1 + P1 + 1 <= a.n1 && p2 >= 1 && p2 + P2 + 1 <= a.n2)
                tma_tensor3d_prefetch_l2(&tmx, (int)((long long)chunk2 * CE * (long long)(sizeof(E) / 8)), (p2 - 1) * RC, p1 - 1);
        }
    }

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

    auto lx = [&](auto U1, auto U2, auto B, int j) {
        return xb[((decltype(U1)::value * (P2 + 2) + decltype(U2)::value) * RC + decltype(B)::value) * CE + 32 * j];
    };
    auto lh = [&](auto V1, auto V2, auto A, auto S) {
        return hb[((decltype(V1)::value * P2 + decltype(V2)::value) * RC + decltype(A)::value) * SWP + decltype(S)::value];
    };
    bool shared_bonds = false;
    if constexpr (st_symmetric<RC>(MK::mask))
        shared_bonds = a.herm && q1 + T1 <= a.n1 && q2 + T2 <= a.n2;      // warp-uniform
    constexpr bool SELF = (MODE == 2 || MODE == 3);
    if (ri) {
        // every value is purely real or purely imaginary (the model's class, MK::imag): 64-bit uniform value
        // loads (one shared-memory wavefront instead of two) and two FMAs per complex element instead of four
        const T* hr = reinterpret_cast<const T*>(sh) + ((w1 * T1) * P2 + w2 * T2) * RC * SWR;
        auto lr = [&](auto V1, auto V2, auto A, auto S) {
            return hr[((decltype(V1)::value * P2 + decltype(V2)::value) * RC + decltype(A)::value) * SWR + decltype(S)::value];
        };
        if (shared_bonds) {
            if constexpr (st_symmetric<RC>(MK::mask)) st_tile_herm<T, RC, MK, T1, T2, CPT, SELF, 1>(acc, g, lx, lr);
        } else {
            st_tile<T, RC, MK, T1, T2, CPT, SELF, 1>(acc, g, lx, lr);
        }
    } else if (shared_bonds) {
        if constexpr (st_symmetric<RC>(MK::mask)) st_tile_herm<T, RC, MK, T1, T2, CPT, SELF, 0>(acc, g, lx, lh);
    } else {
        st_tile<T, RC, MK, T1, T2, CPT, SELF, 0>(acc, g, lx, lh);
    }

    const T2c alpha = cmake<T2c>(a.alpha[0], a.alpha[1]);
    const bool unit_alpha = a.alpha[0] == 1.0 && a.alpha[1] == 0.0;
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
                    if (cj >= nce) continue;
                    const long long e = row * lde + cj;
                    E res;
                    if (unit_alpha) res = acc[v1][v2][aa][j];           // scaling deferred to a later factor (run_factors)
                    else pscale(res, alpha, acc[v1][v2][aa][j]);
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

// ------------------------------------------------------------------------------------------
// k_apply_stencil_stream: streaming variant of the staged kernel.  A CTA owns a patch and a GROUP
// of consecutive column chunks: the value rows of the patch are staged once, the haloed Psi rows
// of chunk k + 2 are in flight (multi-stage TMA pipeline) while chunk k runs out of shared memory.
// One CTA per SM; consecutive CTAs take consecutive patches of the SAME group, so the halo rows
// neighbouring patches share are read by SMs that are at the same chunk at about the same time.
// MEASURED SLOWER than the single-shot kernel above (Haldane 500 x 500: 0.56-0.60 vs 0.69-0.73 of
// the roofline, profiles/stencil_variants_r1.jsonl variants 10-12): with one resident CTA the
// per-chunk CTA barrier is exposed, while three single-shot CTAs per SM overlap each other's copies
// and FMAs for free.  Built only with LM_STENCIL_EXPLORE.
// ------------------------------------------------------------------------------------------
constexpr int ST_STREAM_STAGES = 3;
template <typename T, int RC, typename MK, int T1, int T2, int W1, int W2>
__host__ __device__ constexpr size_t st_stream_smem() {
    return (size_t)ST_STREAM_STAGES * (W1 * T1 + 2) * (W2 * T2 + 2) * RC * 32 * 16
         + (size_t)(W1 * T1) * (W2 * T2) * RC * st_stride<T, RC, MK>() * (2 * sizeof(T));
}

template <typename T, int RC, typename MK, int T1, int T2, int W1, int W2, int MODE>
__global__ void __launch_bounds__(32 * W1 * W2, 1)
k_apply_stencil_stream(const StencilArgs a) {
    using T2c = typename cx2<T>::type;
    using E = typename pack<T>::E;
    constexpr int EC = pack<T>::EC;
    constexpr int NS = ST_STREAM_STAGES;
    constexpr int NT = 32 * W1 * W2, P1 = W1 * T1, P2 = W2 * T2;
    constexpr int HR = (P1 + 2) * (P2 + 2) * RC;
    constexpr int CE = 32;
    constexpr int SWP = st_stride<T, RC, MK>();
    LM_SMEM_DYN(lm_smem);
    E* sx = reinterpret_cast<E*>(lm_smem);                                          // [NS][HR][CE]
    T2c* sh = reinterpret_cast<T2c*>(lm_smem + (size_t)NS * HR * CE * sizeof(E));    // [P1][P2 * RC * SWP]
    LM_SMEM_STATIC __align__(8) unsigned long long bar[NS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned group = blockIdx.x / a.npatch, patch = blockIdx.x - group * a.npatch;
    const int pj1 = (int)(patch / (unsigned)a.np2), pj2 = (int)(patch - (unsigned)pj1 * (unsigned)a.np2);
    const int o1 = pj1 * P1, o2 = pj2 * P2;
    const long long lde = a.ld / EC, nce = lde;
    const unsigned ch0 = group * a.cpg;
    const unsigned ch1 = (ch0 + a.cpg) < a.nchunks ? (ch0 + a.cpg) : a.nchunks;
    if (ch0 >= ch1) return;
    const int vl1 = (a.n1 - o1) < P1 ? (a.n1 - o1) : P1;
    const int vl2 = (a.n2 - o2) < P2 ? (a.n2 - o2) : P2;
    const unsigned hline = ((unsigned)(vl2 * RC * SWP * (int)sizeof(T2c)) + 15u) & ~15u;
    const E* __restrict__ x = (const E*)a.x;
    const T2c* __restrict__ sv = (const T2c*)a.svals;

    if (tid == 0) {
#pragma unroll
        for (int q = 0; q < NS; ++q) mbar_init(&bar[q], 1);
    }
    __syncthreads();
    auto issue = [&](unsigned ch, int stage, bool with_values) {
        const long long c0 = (long long)ch * CE;
        const int cw = (int)((nce - c0) < CE ? (nce - c0) : CE);
        if (tid == 0) mbar_arrive_expect_tx(&bar[stage], (unsigned)(HR * cw * (int)sizeof(E)) + (with_values ? (unsigned)vl1 * hline : 0u));
        for (int r = warp + (NT / 32) * lane; r < HR; r += NT) {       // rows dealt round-robin over the warps
            const int u1 = r / ((P2 + 2) * RC), rem = r - u1 * ((P2 + 2) * RC), u2 = rem / RC, b = rem - u2 * RC;
            const long long row = ((long long)st_wrap(o1 + u1 - 1, a.n1) * a.n2 + st_wrap(o2 + u2 - 1, a.n2)) * RC + b;
            tma_bulk_g2s(sx + ((size_t)stage * HR + r) * CE, x + row * lde + c0, (unsigned)(cw * (int)sizeof(E)), &bar[stage]);
        }
        if (with_values)
            for (int l = NT - 1 - tid; l < vl1; l += NT)
                tma_bulk_g2s(sh + l * (P2 * RC * SWP), sv + ((long long)(o1 + l) * a.n2 + o2) * (RC * SWP), hline, &bar[stage]);
    };

    const int w1 = warp / W2, w2 = warp % W2;
    const int q1 = o1 + w1 * T1, q2 = o2 + w2 * T2;
    const bool active = q1 < a.n1 && q2 < a.n2;
    const T2c* hb = sh + ((w1 * T1) * P2 + w2 * T2) * RC * SWP;
    const T2c g = cmake<T2c>(a.g[0], a.g[1]);
    const T2c alpha = cmake<T2c>(a.alpha[0], a.alpha[1]);
    const T2c beta  = cmake<T2c>(a.beta[0],  a.beta[1]);
    const T2c delta = cmake<T2c>(a.delta[0], a.delta[1]);
    E* y = (E*)a.y;
    const E* z = (const E*)a.z;
    const E* u = (const E*)a.u;

#pragma unroll
    for (int q = 0; q < NS - 1; ++q) if (ch0 + q < ch1) issue(ch0 + q, q, q == 0);
    for (unsigned ch = ch0; ch < ch1; ++ch) {
        const int k = (int)(ch - ch0), stage = k % NS;
        // the stage refilled here was read in iteration k - 1 and released by the barrier below
        if (ch + (NS - 1) < ch1) issue(ch + (NS - 1), (k + NS - 1) % NS, false);
        const long long cj = (long long)ch * CE + lane;
        mbar_wait(&bar[stage], (unsigned)((k / NS) & 1));
        if (active) {
            const E* xb = sx + ((size_t)stage * HR + ((w1 * T1) * (P2 + 2) + w2 * T2) * RC) * CE + lane;
            E acc[T1][T2][RC][1];
#pragma unroll
            for (int v1 = 0; v1 < T1; ++v1)
#pragma unroll
                for (int v2 = 0; v2 < T2; ++v2)
#pragma unroll
                    for (int aa = 0; aa < RC; ++aa) pzero(acc[v1][v2][aa][0]);
            st_tile<T, RC, MK, T1, T2, 1, (MODE == 2 || MODE == 3), 0>(acc, g,
                [&](auto U1, auto U2, auto B, int) {
                    return xb[((decltype(U1)::value * (P2 + 2) + decltype(U2)::value) * RC + decltype(B)::value) * CE];
                },
                [&](auto V1, auto V2, auto A, auto S) {
                    return hb[((decltype(V1)::value * P2 + decltype(V2)::value) * RC + decltype(A)::value) * SWP + decltype(S)::value];
                });
            if (cj < nce) {
#pragma unroll
                for (int v1 = 0; v1 < T1; ++v1)
#pragma unroll
                    for (int v2 = 0; v2 < T2; ++v2) {
                        if (q1 + v1 >= a.n1 || q2 + v2 >= a.n2) continue;      // ragged last tile
#pragma unroll
                        for (int aa = 0; aa < RC; ++aa) {
                            const long long e = (((long long)(q1 + v1) * a.n2 + (q2 + v2)) * RC + aa) * lde + cj;
                            E res;
                            pscale(res, alpha, acc[v1][v2][aa][0]);
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
        __syncthreads();                                    // every warp is done with `stage`
    }
}

// ------------------------------------------------------------------------------------------
// k_observe_stencil: fused localdensity + bond correlators on the same cell-major lattices.
//   dens[i] = sum_c w_c |x[i,c]|^2 ,   G[e] = sum_c w_c x[j,c] conj(x[i,c])  for the upper entry e = (i -> j)
// Every bond has exactly one FORWARD direction (cell offset (0,+1), (+1,*), or a later row of the
// same cell), so a thread that owns a T1 x T2 block of cells of ONE column needs the own rows and
// the forward halo only.  A CTA owns a patch and a GROUP of column chunks: it streams the chunks
// through a multi-stage TMA pipeline (chunks k+1, k+2 land while chunk k is reduced) and keeps its partial
// sums in registers over the whole group - the cross-lane (= cross-column) reduction and the
// atomics run once per group, not once per chunk.  `out` maps (row, forward slot) to the ELL
// entry that carries the pair: e >= 0 as is, e <= -2 the conjugate goes to entry -2 - e (the
// forward neighbour sits across a periodic boundary), -1 no such bond on this row.
// ------------------------------------------------------------------------------------------
template <int RC> __host__ __device__ constexpr bool st_is_fwd(const st_mask_t& m, int o, int a, int b) {
    return st_bit<RC>(m, o, a, b) && (o > 4 || (o == 4 && b > a));
}
template <int RC> __host__ __device__ constexpr int st_fslot(const st_mask_t& m, int o, int a, int b) {
    int s = 0;
    for (int oo = 4; oo < 9; ++oo)
        for (int bb = 0; bb < RC; ++bb) {
            if (oo == o && bb == b) return s;
            if (st_is_fwd<RC>(m, oo, a, bb)) ++s;
        }
    return s;
}
template <int RC> __host__ __device__ constexpr int st_nfwd(const st_mask_t& m) {
    int w = 1;
    for (int a = 0; a < RC; ++a) {
        int s = 0;
        for (int o = 4; o < 9; ++o) for (int b = 0; b < RC; ++b) if (st_is_fwd<RC>(m, o, a, b)) ++s;
        if (s > w) w = s;
    }
    return w;
}
// is in row b of staged cell (l1, u2) - l1 in [0, T1], u2 in [0, T2 + 2) - a forward neighbour of the tile?
template <int RC, int T1, int T2> __host__ __device__ constexpr bool st_fwd_needed(const st_mask_t& m, int l1, int u2, int b) {
    for (int o = 4; o < 9; ++o)
        for (int a = 0; a < RC; ++a)
            if (st_is_fwd<RC>(m, o, a, b)) {
                const int v1 = l1 - (o / 3 - 1), v2 = u2 - 1 - (o % 3 - 1);
                if (v1 >= 0 && v1 < T1 && v2 >= 0 && v2 < T2) return true;
            }
    return false;
}

struct StencilObsArgs {
    int n1, n2, np2;
    long long M, ld;                // logical columns, leading dimension (complex elements)
    const void* x; const double* w; // w may be null (all ones)
    const int* out;                 // [N][NF]
    double* dens; double2* G;
    unsigned ngroups, cpg, nchunks; // column groups per patch, chunks per group, chunks in total
    unsigned npatch;                // patches (grid = ngroups x npatch, patch fastest)
    int tmap;                       // 1: patches whose forward-haloed block lies inside the lattice stage it with ONE tensor-map copy per chunk
};

template <typename T> struct st_unpack;
template <> struct st_unpack<double> {
    static __device__ __forceinline__ void get(const double2& e, int, double& re, double& im) { re = e.x; im = e.y; }
};
template <> struct st_unpack<float> {
    static __device__ __forceinline__ void get(const float4& e, int k, double& re, double& im) {
        re = k ? (double)e.z : (double)e.x; im = k ? (double)e.w : (double)e.y;
    }
};
__device__ __forceinline__ double st_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

constexpr int ST_OBS_STAGES = 3;      // TMA pipeline depth: ~2 staged chunks in flight per CTA cover the HBM latency
template <typename T, int RC, int T1, int T2, int W1, int W2>
__host__ __device__ constexpr size_t st_obs_smem() { return (size_t)ST_OBS_STAGES * (W1 * T1 + 1) * (W2 * T2 + 2) * RC * 32 * 16; }

template <typename T, int RC, typename MK, int T1, int T2, int W1, int W2>
__global__ void __launch_bounds__(32 * W1 * W2, 65536 / (128 * 32 * W1 * W2) < 1 ? 1 : 65536 / (128 * 32 * W1 * W2))
k_observe_stencil(const StencilObsArgs a, const LM_GRID_CONSTANT CUtensorMap tmx) {
    constexpr int NS = ST_OBS_STAGES;
    using E = typename pack<T>::E;
    constexpr int EC = pack<T>::EC;
    constexpr int NT = 32 * W1 * W2, P1 = W1 * T1, P2 = W2 * T2;
    constexpr int L2 = P2 + 2;
    constexpr int HR = (P1 + 1) * L2 * RC;                  // staged rows: own + forward cell lines
    constexpr int CE = 32;
    constexpr int NF = st_nfwd<RC>(MK::mask);
    LM_SMEM_DYN(lm_smem);
    E* sx = reinterpret_cast<E*>(lm_smem);                  // [NS][HR][CE]
    LM_SMEM_STATIC __align__(8) unsigned long long bar[NS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // patch fastest: the CTAs of one column group sweep the lattice together, so the forward-halo rows that
    // neighbouring patches share are read while they are still in L2 (group-fastest order with one long group
    // per patch re-read them from HBM: 1.26x the algorithmic traffic at M = 4096, profiles/r2/bench_default_ncu_r2.md)
    const unsigned group = blockIdx.x / a.npatch, patch = blockIdx.x - group * a.npatch;
    const int pj1 = (int)(patch / (unsigned)a.np2), pj2 = (int)(patch - (unsigned)pj1 * (unsigned)a.np2);
    const int o1 = pj1 * P1, o2 = pj2 * P2;
    const long long lde = a.ld / EC;
    const E* __restrict__ x = (const E*)a.x;
    const unsigned ch0 = group * a.cpg;
    const unsigned ch1 = (ch0 + a.cpg) < a.nchunks ? (ch0 + a.cpg) : a.nchunks;
    if (ch0 >= ch1) return;

    if (tid == 0) {
#pragma unroll
        for (int q = 0; q < NS; ++q) mbar_init(&bar[q], 1);
    }
    __syncthreads();
    // a patch whose forward-haloed block [o1, o1 + P1] x [o2 - 1, o2 + P2] lies inside the lattice is one
    // dense box of the [n1][n2 RC][columns] view of x: ONE tensor-map copy per chunk (SASS UTMALDG).
    // (A per-thread cp.async.bulk is serialised by the compiler into a warp-wide loop of uniform-register
    // broadcasts - ~10 instructions per copy, a fifth of this kernel's instructions in round 1.)
    const bool boxed = a.tmap && o1 + P1 + 1 <= a.n1 && o2 >= 1 && o2 + P2 + 1 <= a.n2;
    auto issue = [&](unsigned ch, int stage) {              // arm the stage barrier, then the box or one bulk copy per staged row
        const long long c0 = (long long)ch * CE;
        const int cw = (int)((lde - c0) < CE ? (lde - c0) : CE);
        if (boxed) {
            if (tid == 0) {
                mbar_arrive_expect_tx(&bar[stage], (unsigned)(HR * CE * (int)sizeof(E)));
                tma_tensor3d_g2s(sx + (size_t)stage * HR * CE, &tmx, (int)(c0 * (long long)(sizeof(E) / 8)), (o2 - 1) * RC, o1, &bar[stage]);
            }
            return;
        }
        if (tid == 0) mbar_arrive_expect_tx(&bar[stage], (unsigned)(HR * cw * (int)sizeof(E)));
        // rows dealt round-robin over the WARPS (row = warp + NW * lane): every warp issues the same
        // few copies, none of them reaches the closing barrier late
        for (int r = warp + (NT / 32) * lane; r < HR; r += NT) {
            const int l1 = r / (L2 * RC), rem = r - l1 * (L2 * RC), u2 = rem / RC, b = rem - u2 * RC;
            const long long row = ((long long)st_wrap(o1 + l1, a.n1) * a.n2 + st_wrap(o2 + u2 - 1, a.n2)) * RC + b;
            tma_bulk_g2s(sx + ((size_t)stage * HR + r) * CE, x + row * lde + c0, (unsigned)(cw * (int)sizeof(E)), &bar[stage]);
        }
    };

    const int w1 = warp / W2, w2 = warp % W2;
    const int q1 = o1 + w1 * T1, q2 = o2 + w2 * T2;
    const bool active = q1 < a.n1 && q2 < a.n2;
    double dens[T1][T2][RC];
    double gr[T1][T2][RC][NF], gi[T1][T2][RC][NF];
#pragma unroll
    for (int v1 = 0; v1 < T1; ++v1)
#pragma unroll
        for (int v2 = 0; v2 < T2; ++v2)
#pragma unroll
            for (int aa = 0; aa < RC; ++aa) {
                dens[v1][v2][aa] = 0.0;
#pragma unroll
                for (int f = 0; f < NF; ++f) { gr[v1][v2][aa][f] = 0.0; gi[v1][v2][aa][f] = 0.0; }
            }

#pragma unroll
    for (int q = 0; q < NS - 1; ++q) if (ch0 + q < ch1) issue(ch0 + q, q);
    for (unsigned ch = ch0; ch < ch1; ++ch) {
        const int k = (int)(ch - ch0), stage = k % NS;
        // the stage refilled here was read in iteration k - 1 and released by the barrier below
        // (a warp-rotating producer with per-stage "empty" mbarriers instead of the CTA barrier was
        // measured 1.6x slower on Haldane 500 x 500)
        if (ch + (NS - 1) < ch1) issue(ch + (NS - 1), (k + NS - 1) % NS);
        const long long c0 = (long long)ch * CE;
        const bool lane_on = active && (c0 + lane) < lde;
        double wl[EC];
#pragma unroll
        for (int e = 0; e < EC; ++e) {
            const long long col = (c0 + lane) * EC + e;
            wl[e] = (col < a.M) ? (a.w ? a.w[col] : 1.0) : 0.0;
        }
        mbar_wait(&bar[stage], (unsigned)((k / NS) & 1));
        if (lane_on) {
            const E* xb = sx + ((size_t)stage * HR + ((w1 * T1) * L2 + w2 * T2) * RC) * CE + lane;
            double ar[T1][T2][RC][EC], ai[T1][T2][RC][EC];
            st_for<T1>([&](auto V1) { st_for<T2>([&](auto V2) { st_for<RC>([&](auto A) {
                constexpr int v1 = decltype(V1)::value, v2 = decltype(V2)::value, aa = decltype(A)::value;
                const E xe = xb[((v1 * L2 + v2 + 1) * RC + aa) * CE];
#pragma unroll
                for (int e = 0; e < EC; ++e) {
                    double re, im; st_unpack<T>::get(xe, e, re, im);
                    ar[v1][v2][aa][e] = wl[e] * re; ai[v1][v2][aa][e] = wl[e] * im;
                    dens[v1][v2][aa] = fma(ar[v1][v2][aa][e], re, dens[v1][v2][aa]);
                    dens[v1][v2][aa] = fma(ai[v1][v2][aa][e], im, dens[v1][v2][aa]);
                }
            }); }); });
            st_for<T1 + 1>([&](auto LL1) { st_for<T2 + 2>([&](auto U2) { st_for<RC>([&](auto B) {
                constexpr int l1 = decltype(LL1)::value, u2 = decltype(U2)::value, b = decltype(B)::value;
                if constexpr (st_fwd_needed<RC, T1, T2>(MK::mask, l1, u2, b)) {
                    const E xe = xb[((l1 * L2 + u2) * RC + b) * CE];
                    st_for<5>([&](auto OO) { st_for<RC>([&](auto A) {
                        constexpr int o = decltype(OO)::value + 4, aa = decltype(A)::value;
                        constexpr int v1 = l1 - (o / 3 - 1), v2 = u2 - 1 - (o % 3 - 1);
                        if constexpr (st_is_fwd<RC>(MK::mask, o, aa, b) && v1 >= 0 && v1 < T1 && v2 >= 0 && v2 < T2) {
                            constexpr int f = st_fslot<RC>(MK::mask, o, aa, b);
#pragma unroll
                            for (int e = 0; e < EC; ++e) {
                                double re, im; st_unpack<T>::get(xe, e, re, im);
                                // x_j * conj(x_i) * w
                                gr[v1][v2][aa][f] = fma(re, ar[v1][v2][aa][e], gr[v1][v2][aa][f]);
                                gr[v1][v2][aa][f] = fma(im, ai[v1][v2][aa][e], gr[v1][v2][aa][f]);
                                gi[v1][v2][aa][f] = fma(im, ar[v1][v2][aa][e], gi[v1][v2][aa][f]);
                                gi[v1][v2][aa][f] = fma(-re, ai[v1][v2][aa][e], gi[v1][v2][aa][f]);
                            }
                        }
                    }); });
                }
            }); }); });
        }
        __syncthreads();                                    // every warp is done with `stage`
    }
    if (!active) return;
    // Fold the 32 columns of the lane dimension by recursive halving: at distance 16, 8, ... 1 every lane
    // hands one half of its remaining sums to its partner and keeps the other - NV - 1 + 5 exchanges
    // per lane instead of 5 NV for a butterfly per sum - and ends up with the complete sums number
    // 2 lane and 2 lane + 1, which it adds to the output (one atomic per sum and group).
    constexpr int PER = 1 + 2 * NF;                          // per row: dens, NF x Re G, NF x Im G
    constexpr int NV = T1 * T2 * RC * PER;
    static_assert(NV <= 64, "k_observe_stencil: more than 64 partial sums per thread");
    double v[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = 0.0;
#pragma unroll
    for (int v1 = 0; v1 < T1; ++v1)
#pragma unroll
        for (int v2 = 0; v2 < T2; ++v2)
#pragma unroll
            for (int aa = 0; aa < RC; ++aa) {
                const int base = ((v1 * T2 + v2) * RC + aa) * PER;
                v[base] = dens[v1][v2][aa];
#pragma unroll
                for (int f = 0; f < NF; ++f) { v[base + 1 + f] = gr[v1][v2][aa][f]; v[base + 1 + NF + f] = gi[v1][v2][aa][f]; }
            }
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int off = 16 >> s, n = 32 >> s;                // n sums stay with each lane after this round (of 2 n)
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n; ++i) {
            if (i < NV) {                                    // compile-time: slot i only ever holds sums >= i
                const double send = up ? v[i] : v[i + n];
                const double keep = up ? v[i + n] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
        }
    }
    // round s (n = 32 >> s) keeps in slot i what slot i + n * [lane bit (4 - s)] held: after the last round
    // slot k in {0, 1} of a lane holds sum number 2 lane + k
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int j = 2 * lane + k;
        if (j >= NV) continue;
        const int cell = j / (RC * PER), rem = j - cell * (RC * PER), aa = rem / PER, q = rem - aa * PER;
        const int v1 = cell / T2, v2 = cell - v1 * T2;
        if ((q1 + v1) >= a.n1 || (q2 + v2) >= a.n2) continue;
        const long long row = ((long long)(q1 + v1) * a.n2 + (q2 + v2)) * RC + aa;
        const double val = k ? v[1] : v[0];
        if (q == 0) { atomicAdd(a.dens + row, val); continue; }
        const int f = (q - 1) % NF; const bool imag = (q - 1) >= NF;
        const int oe = a.out[row * NF + f];
        if (oe == -1) continue;
        double* g = reinterpret_cast<double*>(a.G + (oe >= 0 ? oe : -2 - oe));
        if (!imag) atomicAdd(g, val); else atomicAdd(g + 1, oe >= 0 ? val : -val);
    }
}

// ---- compiled patterns (stencil.cu registry; one translation unit per pattern) ----
// mask: the stored entries.  imag: the entries that are PURELY IMAGINARY in the pattern's "real / imaginary"
// value class (every other entry purely real) - the class of the reference's own models without a magnetic
// field: real hoppings and on-site terms, `im * t2` second-neighbour hops of `haldane`
// (src/zoo/models.jl:164-170), the `-im/2` orbital-flip x-hops of `qwz` (src/zoo/models.jl:130-136).  Whether
// the CURRENT values are in the class is decided on the device whenever they change (k_gather_real, api.cu);
// anything else (Peierls phases, twists) runs the general complex path of the same kernel.
#define LM_ST_MASK0 0xbaull
#define LM_ST_MASK1 0x1ffull
#define LM_ST_MASK2 0x404f2020ull
#define LM_ST_MASK3 0xf0fff0f0ull
#define LM_ST_MASK4 0xd9dfb9b0ull
#define LM_ST_MASK5 0xfffffffffull
#define LM_ST_IMAG3 0x60000060ull
#define LM_ST_IMAG4 0x99909990ull
template <int ID> struct StPat;
template <> struct StPat<0> { static constexpr int rc = 1; static constexpr st_mask_t mask = {{LM_ST_MASK0, 0, 0, 0}}, imag = {{0, 0, 0, 0}}; };
template <> struct StPat<1> { static constexpr int rc = 1; static constexpr st_mask_t mask = {{LM_ST_MASK1, 0, 0, 0}}, imag = {{0, 0, 0, 0}}; };
template <> struct StPat<2> { static constexpr int rc = 2; static constexpr st_mask_t mask = {{LM_ST_MASK2, 0, 0, 0}}, imag = {{0, 0, 0, 0}}; };
template <> struct StPat<3> { static constexpr int rc = 2; static constexpr st_mask_t mask = {{LM_ST_MASK3, 0, 0, 0}}, imag = {{LM_ST_IMAG3, 0, 0, 0}}; };
template <> struct StPat<4> { static constexpr int rc = 2; static constexpr st_mask_t mask = {{LM_ST_MASK4, 0, 0, 0}}, imag = {{LM_ST_IMAG4, 0, 0, 0}}; };
template <> struct StPat<5> { static constexpr int rc = 2; static constexpr st_mask_t mask = {{LM_ST_MASK5, 0, 0, 0}}, imag = {{0, 0, 0, 0}}; };
// three and four rows per cell: the masks span two words
//   6  kagome NN + on-site (src/zoo/lattices.jl:209, NearestNeighbor(1))       7  kagome NN + second neighbours + on-site
//   8  Kane-Mele (src/zoo/models.jl:188-194): spin-diagonal honeycomb NN + `im t2 sigma_z` second neighbours + on-site diagonal
template <> struct StPat<6> { static constexpr int rc = 3; static constexpr st_mask_t mask = {{0x8081ff022000400ull, 0x4ull, 0, 0}}, imag = {{0, 0, 0, 0}}; };
template <> struct StPat<7> { static constexpr int rc = 3; static constexpr st_mask_t mask = {{0xb191ff133090c00ull, 0x34ull, 0, 0}}, imag = {{0, 0, 0, 0}}; };
template <> struct StPat<8> { static constexpr int rc = 4; static constexpr st_mask_t mask = {{0x84a5842184a50000ull, 0xa5218421a521a5a5ull, 0, 0}},
                                                                                      imag = {{0x8421842184210000ull, 0x8421842184210000ull, 0, 0}}; };
//   9  QWZ as the reference builds it (src/zoo/models.jl:130-136): the on-site term `sigma_z m` is diagonal, 9 instead of 10 entries per row
template <> struct StPat<9> { static constexpr int rc = 2; static constexpr st_mask_t mask = {{0xf0f9f0f0ull, 0, 0, 0}}, imag = {{LM_ST_IMAG3, 0, 0, 0}}; };
constexpr int LM_ST_NPAT = 10;

#ifndef __CUDACC_RTC__
// ---- host-visible registry (stencil.cu) ----
struct StencilDesc { int rc; st_mask_t mask, imag; int sw; const char* name; };
int stencil_count();
const StencilDesc& stencil_desc(int id);
int stencil_find(int rc, const st_mask_t& mask);                  // smallest compiled superset, -1 if none
int stencil_stride(int id, bool c64);                      // value-slot stride (complex64 rows are padded to even)
int stencil_rstride(int id, bool c64);                     // scalars per row of the real / imaginary class copy (st_rstride)
int stencil_diag_slot(int id, int a);                      // slot of the diagonal entry of in-cell row a
int stencil_num_variants();
// patch size in cells, lane elements per thread and kernel family (staged = TMA) of a variant
void stencil_variant_shape(int variant, int rc, int* P1, int* P2, int* cpt, int* staged);
int stencil_resident_ctas(int id, int variant, bool c64);  // CTAs of the staged kernel resident per SM
// launches; returns 0 on success, -1 if (id, variant, mode) is not compiled, -2 on a CUDA error
int stencil_launch(int id, int variant, bool c64, int mode, const StencilArgs& a, const CUtensorMap& tmx, dim3 grid, cudaStream_t s);
// fused observables: patch size / forward-slot count of the compiled kernel, and its launch
void stencil_obs_shape(int id, int* P1, int* P2, int* nf);
int stencil_observe(int id, bool c64, const StencilObsArgs& a, const CUtensorMap& tmx, unsigned grid, cudaStream_t s);
// ---- run-time compiled patterns (stencil_rtc.cu): ids >= LM_ST_RTC_BASE; every registry function above accepts them ----
constexpr int LM_ST_RTC_BASE = 1000;
bool stencil_rtc_available();                                                       // NVRTC + driver entry points found, LM_STENCIL_RTC != 0
int stencil_rtc_register(int rc, const st_mask_t& mask, const st_mask_t& imag);     // id of the pattern (registered once), -1 if unavailable
const StencilDesc* stencil_rtc_desc(int id);
int stencil_rtc_warm(int id, bool c64);                                             // compile the step kernels now (0 on success)
const char* stencil_rtc_error();
void stencil_rtc_apply_shape(int rc, int* t1, int* t2, int* w1, int* w2);
bool stencil_rtc_obs_shape(int rc, int nf, int* t1, int* t2, int* w1, int* w2);
int stencil_rtc_nfwd(int rc, const st_mask_t& m);
int stencil_rtc_launch(int id, bool c64, int mode, const StencilArgs& a, const CUtensorMap& tmx, dim3 grid, cudaStream_t s);
int stencil_rtc_observe(int id, bool c64, const StencilObsArgs& a, const CUtensorMap& tmx, unsigned grid, cudaStream_t s);
#endif  // __CUDACC_RTC__

}  // namespace lm
