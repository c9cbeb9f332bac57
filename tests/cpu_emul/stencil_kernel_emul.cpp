// CPU execution of the WHOLE stencil kernels of csrc/stencil.cuh - k_apply_stencil_tma (the default
// SpMM / propagator-factor kernel), k_apply_stencil (direct loads) and k_observe_stencil (fused
// localdensity + bond correlators) - against independently written reference loops.  Every CUDA
// thread of a CTA is a real OS thread (shim/cuda_runtime.h: real __syncthreads, warp-shuffle
// mailboxes, atomics, the mbarrier phase rule and 16-byte checked bulk copies), so the staging
// index math, the ragged patch / chunk handling, the tensor-map box of interior patches, the shared
// value loads of Hermitian operators and the periodic wrap are exercised exactly as written for the device.
// TEST INFRASTRUCTURE ONLY.  Prints "OK <n checks>" or the first mismatch; exit code 0 / 1.
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "stencil.cuh"

using namespace lm;
typedef std::complex<double> zc;
static std::mt19937_64 rng(987654321);
static double rnd() { return std::uniform_real_distribution<double>(-1.0, 1.0)(rng); }
static long long nchecks = 0;

static int wrap(int c, int n) { c %= n; return c < 0 ? c + n : c; }

template <typename T> static zc get_el(const typename pack<T>::E* p, long long lde, long long row, long long col) {
    constexpr int EC = pack<T>::EC;
    const typename pack<T>::E& e = p[row * lde + col / EC];
    if constexpr (EC == 1) return zc(e.x, e.y);
    else return (col & 1) ? zc(e.z, e.w) : zc(e.x, e.y);
}
template <typename T> static void set_el(typename pack<T>::E* p, long long lde, long long row, long long col, zc v) {
    constexpr int EC = pack<T>::EC;
    typename pack<T>::E& e = p[row * lde + col / EC];
    if constexpr (EC == 1) { e.x = v.real(); e.y = v.imag(); }
    else if (col & 1) { e.z = (float)v.real(); e.w = (float)v.imag(); }
    else { e.x = (float)v.real(); e.y = (float)v.imag(); }
}

// ---------------------------------------------------------------------------------------------
// y = alpha (H x + g x) + beta z + delta u   on an n1 x n2 lattice
// STAGED: 1 = k_apply_stencil_tma, 0 = k_apply_stencil
// flags: bit 0 = Hermitian values + shared value loads (StencilArgs.herm), bit 1 = tensor-map boxes (StencilArgs.tmap),
//        bit 2 = values in the pattern's real / imaginary class, read as scalars (StencilArgs.sreal / ri_flag),
//        bit 3 = class copy supplied but the device flag says "not in the class": the complex values must be used
// ---------------------------------------------------------------------------------------------
template <typename T, int RC, typename MK, int T1, int T2, int W1, int W2, int CPT, int MODE, int STAGED>
static bool check_apply(const char* name, int n1, int n2, bool periodic, long long ld, unsigned cps, int flags) {
    const long long c_off = 0, nc = ld;
    const bool herm = flags & 1, tmap = flags & 2, ri = flags & 4, ri_off = flags & 8;
    using E = typename pack<T>::E;
    using T2c = typename cx2<T>::type;
    constexpr int EC = pack<T>::EC;
    constexpr int SWP = st_stride<T, RC, MK>();
    constexpr int SWR = st_rstride<T, RC, MK>();
    constexpr int P1 = W1 * T1, P2 = W2 * T2;
    const long long N = (long long)n1 * n2 * RC, lde = ld / EC;
    std::vector<E> x(N * lde), y(N * lde), z(N * lde), u(N * lde);
    std::vector<T> sr(N * SWR + 64, (T)12345);             // class scalars (padding slots hold garbage on purpose)
    std::vector<T2c> sv(N * SWP);
    std::vector<zc> svz(N * SWP, zc(0, 0));
    for (long long r = 0; r < N; ++r)
        for (long long c = 0; c < ld; ++c) {
            set_el<T>(x.data(), lde, r, c, zc(rnd(), rnd()));
            set_el<T>(z.data(), lde, r, c, zc(rnd(), rnd()));
            set_el<T>(u.data(), lde, r, c, zc(rnd(), rnd()));
            set_el<T>(y.data(), lde, r, c, zc(777.0, -777.0));
        }
    for (int j1 = 0; j1 < n1; ++j1) for (int j2 = 0; j2 < n2; ++j2) for (int a = 0; a < RC; ++a) {
        const long long row = ((long long)j1 * n2 + j2) * RC + a;
        int s = 0;
        for (int o = 0; o < 9; ++o) for (int b = 0; b < RC; ++b) {
            if (!(st_bit<RC>(MK::mask, o, a, b))) continue;
            const int k1 = j1 + o / 3 - 1, k2 = j2 + o % 3 - 1;
            const bool inside = k1 >= 0 && k1 < n1 && k2 >= 0 && k2 < n2;
            zc v = (inside || periodic) ? zc(rnd(), rnd()) : zc(0, 0);   // entries absent at open boundaries hold 0
            if (ri) v = st_bit<RC>(MK::imag, o, a, b) ? zc(0, v.imag()) : zc(v.real(), 0);
            svz[row * SWP + s] = v;
            ++s;
        }
    }
    if (herm) {
        // H[q, p] = conj(H[p, q]): the mirror of slot (o, a, b) of row p is slot (8 - o, b, a) of the neighbour row
        for (int j1 = 0; j1 < n1; ++j1) for (int j2 = 0; j2 < n2; ++j2) for (int a = 0; a < RC; ++a) {
            const long long row = ((long long)j1 * n2 + j2) * RC + a;
            for (int o = 0; o < 9; ++o) for (int b = 0; b < RC; ++b) {
                if (!(st_bit<RC>(MK::mask, o, a, b))) continue;
                const int k1 = j1 + o / 3 - 1, k2 = j2 + o % 3 - 1;
                const bool inside = k1 >= 0 && k1 < n1 && k2 >= 0 && k2 < n2;
                zc& v = svz[row * SWP + st_slot<RC>(MK::mask, o, a, b)];
                if (o == 4 && a == b) { v = zc(v.real(), 0.0); continue; }
                if (!(inside || periodic)) continue;
                if (!(o > 4 || (o == 4 && b > a))) continue;                  // forward entries define their mirrors
                const long long nb = ((long long)wrap(k1, n1) * n2 + wrap(k2, n2)) * RC + b;
                svz[nb * SWP + st_slot<RC>(MK::mask, 8 - o, b, a)] = std::conj(v);
            }
        }
    }
    for (long long i = 0; i < N * SWP; ++i) sv[i] = cmake<T2c>(svz[i].real(), svz[i].imag());
    int ri_flag = ri_off ? 0 : 1;
    if (ri || ri_off)
        for (long long row = 0; row < N; ++row) {
            const int a = (int)(row % RC);
            for (int o = 0; o < 9; ++o) for (int b = 0; b < RC; ++b) {
                if (!st_bit<RC>(MK::mask, o, a, b)) continue;
                const int sl = st_slot<RC>(MK::mask, o, a, b);
                const zc v = svz[row * SWP + sl];
                sr[row * SWR + sl] = (T)(st_bit<RC>(MK::imag, o, a, b) ? v.imag() : v.real());
            }
        }

    const zc alpha(rnd(), rnd()), g(rnd(), rnd()), beta(rnd(), rnd()), delta(rnd(), rnd());
    StencilArgs a;
    a.svals = sv.data(); a.sreal = (ri || ri_off) ? sr.data() : nullptr; a.ri_flag = (ri || ri_off) ? &ri_flag : nullptr; a.n1 = n1; a.n2 = n2; a.ld = ld; a.pdl = 0; a.herm = herm ? 1 : 0; a.tmap = tmap ? 1 : 0; a.pf = 5;
    a.x = x.data() + c_off / EC; a.y = y.data() + c_off / EC;
    a.z = (MODE == 1 || MODE == 2) ? z.data() + c_off / EC : nullptr;
    a.u = (MODE == 2) ? u.data() + c_off / EC : nullptr;
    a.alpha[0] = alpha.real(); a.alpha[1] = alpha.imag(); a.g[0] = g.real(); a.g[1] = g.imag();
    a.beta[0] = beta.real(); a.beta[1] = beta.imag(); a.delta[0] = delta.real(); a.delta[1] = delta.imag();
    // the launch geometry of api.cu apply_stencil
    const int CT = 32 * CPT * EC;
    const long long nchunks = (nc + CT - 1) / CT;
    const long long np1 = (n1 + P1 - 1) / P1, np2 = (n2 + P2 - 1) / P2;
    a.np2 = (int)np2;
    long long c = std::max<long long>(1, std::min<long long>(cps, nchunks));
    const long long strips = (nchunks + c - 1) / c;
    c = (nchunks + strips - 1) / strips;
    a.cps = (unsigned)c; a.nchunks = (unsigned)nchunks;
    a.ngroups = 1; a.cpg = (unsigned)nchunks; a.npatch = (unsigned)(np1 * np2);
    dim3 grid((unsigned)(np1 * np2 * c), (unsigned)strips);
    if constexpr (STAGED == 1) {
        static_assert(st_tma_smem<T, RC, MK, T1, T2, W1, W2, CPT>() <= 227 * 1024, "patch does not fit shared memory");
        CUtensorMap tmx{x.data(), {(unsigned long long)ld * (2 * sizeof(T) / 8), (unsigned long long)n2 * RC, (unsigned long long)n1},
                        {(unsigned long long)ld * 2 * sizeof(T), (unsigned long long)n2 * RC * ld * 2 * sizeof(T)},
                        {(unsigned)(32 * CPT * 2), (unsigned)((P2 + 2) * RC), (unsigned)(P1 + 2)}, tmap ? 1 : 0};
        lm_emul::run_grid(grid, dim3(32 * W1 * W2), [&] { k_apply_stencil_tma<T, RC, MK, T1, T2, W1, W2, CPT, MODE>(a, tmx); });
    } else {
        lm_emul::launch(k_apply_stencil<T, RC, MK, T1, T2, W1, W2, CPT, MODE>, grid, 32 * W1 * W2, a);
    }

    const double tol = sizeof(T) == 8 ? 1e-12 : 5e-5;
    for (int j1 = 0; j1 < n1; ++j1) for (int j2 = 0; j2 < n2; ++j2) for (int aa = 0; aa < RC; ++aa) {
        const long long row = ((long long)j1 * n2 + j2) * RC + aa;
        for (long long col = 0; col < ld; ++col) {
            const zc got = get_el<T>(y.data(), lde, row, col);
            if (col < c_off || col >= c_off + nc) {
                if (got != zc(777.0, -777.0)) { printf("FAIL %s: column %lld outside the window [%lld, %lld) was written (row %lld)\n", name, col, c_off, c_off + nc, row); return false; }
                continue;
            }
            zc hx(0, 0);
            int s = 0;
            for (int o = 0; o < 9; ++o) for (int b = 0; b < RC; ++b) {
                if (!(st_bit<RC>(MK::mask, o, aa, b))) continue;
                const long long nb = ((long long)wrap(j1 + o / 3 - 1, n1) * n2 + wrap(j2 + o % 3 - 1, n2)) * RC + b;
                hx += svz[row * SWP + s] * get_el<T>(x.data(), lde, nb, col);
                ++s;
            }
            zc ref;
            if (MODE == 2 || MODE == 3) hx += g * get_el<T>(x.data(), lde, row, col);
            ref = alpha * hx;
            if (MODE == 1 || MODE == 2) ref += beta * get_el<T>(z.data(), lde, row, col);
            if (MODE == 2) ref += delta * get_el<T>(u.data(), lde, row, col);
            if (!(std::abs(got - ref) <= tol * (1.0 + std::abs(ref)))) {
                printf("FAIL %s: %dx%d %s ld=%lld window [%lld,+%lld) mode %d: row %lld col %lld got (%g,%g) want (%g,%g)\n", name, n1, n2,
                       periodic ? "periodic" : "open", ld, c_off, nc, MODE, row, col, got.real(), got.imag(), ref.real(), ref.imag());
                return false;
            }
            ++nchecks;
        }
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// dens[i] = sum_c w_c |x[i,c]|^2,  G[e] = sum_c w_c x[j,c] conj(x[i,c]) over the forward entries
// ---------------------------------------------------------------------------------------------
template <typename T, int RC, typename MK, int T1, int T2, int W1, int W2>
static bool check_observe(const char* name, int n1, int n2, bool periodic, long long M, long long ld, bool weights, unsigned cpg_req, bool tmap = false) {
    using E = typename pack<T>::E;
    constexpr int EC = pack<T>::EC;
    constexpr int NF = st_nfwd<RC>(MK::mask);
    constexpr int P1 = W1 * T1, P2 = W2 * T2;
    static_assert(st_obs_smem<T, RC, T1, T2, W1, W2>() <= 227 * 1024, "observables patch does not fit shared memory");
    const long long N = (long long)n1 * n2 * RC, lde = ld / EC;
    std::vector<E> x(N * lde);
    // padding columns [M, ld) hold garbage on purpose: the kernel must mask them through the weights
    for (long long r = 0; r < N; ++r) for (long long c = 0; c < ld; ++c) set_el<T>(x.data(), lde, r, c, zc(rnd(), rnd()));
    std::vector<double> w(M);
    for (auto& v : w) v = 0.5 + 0.5 * rnd();
    std::vector<int> out(N * NF, -1);
    std::vector<double> dens(N, 0.0), dens_ref(N, 0.0);
    std::vector<double2> G(N * NF, double2{0, 0});
    std::vector<zc> G_ref(N * NF, zc(0, 0));
    for (int j1 = 0; j1 < n1; ++j1) for (int j2 = 0; j2 < n2; ++j2) for (int a = 0; a < RC; ++a) {
        const long long row = ((long long)j1 * n2 + j2) * RC + a;
        for (long long c = 0; c < M; ++c) dens_ref[row] += (weights ? w[c] : 1.0) * std::norm(get_el<T>(x.data(), lde, row, c));
        int f = 0;
        for (int o = 4; o < 9; ++o) for (int b = 0; b < RC; ++b) {
            if (!(st_bit<RC>(MK::mask, o, a, b)) || !(o > 4 || b > a)) continue;
            const int k1 = j1 + o / 3 - 1, k2 = j2 + o % 3 - 1;
            const bool inside = k1 >= 0 && k1 < n1 && k2 >= 0 && k2 < n2;
            const long long e = row * NF + f;
            if (inside || periodic) {
                const long long nb = ((long long)wrap(k1, n1) * n2 + wrap(k2, n2)) * RC + b;
                zc acc(0, 0);
                for (long long c = 0; c < M; ++c) acc += (weights ? w[c] : 1.0) * get_el<T>(x.data(), lde, nb, c) * std::conj(get_el<T>(x.data(), lde, row, c));
                // across a periodic boundary the kernel hands the CONJUGATE to the entry it is told (here: the same slot)
                out[e] = inside ? (int)e : -2 - (int)e;
                G_ref[e] = inside ? acc : std::conj(acc);
            }
            ++f;
        }
    }
    StencilObsArgs a;
    a.n1 = n1; a.n2 = n2; a.M = M; a.ld = ld; a.x = x.data(); a.w = weights ? w.data() : nullptr;
    a.out = out.data(); a.dens = dens.data(); a.G = G.data();
    const long long np1 = (n1 + P1 - 1) / P1, np2 = (n2 + P2 - 1) / P2;
    a.np2 = (int)np2;
    const long long nchunks = (lde + 31) / 32;
    const long long cpg = std::max<long long>(1, std::min<long long>(cpg_req, nchunks));
    const long long ngroups = (nchunks + cpg - 1) / cpg;
    a.ngroups = (unsigned)ngroups; a.cpg = (unsigned)cpg; a.nchunks = (unsigned)nchunks; a.npatch = (unsigned)(np1 * np2);
    a.tmap = tmap ? 1 : 0;
    CUtensorMap tmx{x.data(), {(unsigned long long)ld * (2 * sizeof(T) / 8), (unsigned long long)n2 * RC, (unsigned long long)n1},
                    {(unsigned long long)ld * 2 * sizeof(T), (unsigned long long)n2 * RC * ld * 2 * sizeof(T)},
                    {64u, (unsigned)((P2 + 2) * RC), (unsigned)(P1 + 1)}, tmap ? 1 : 0};
    lm_emul::run_grid(dim3((unsigned)(np1 * np2 * ngroups)), dim3(32 * W1 * W2), [&] { k_observe_stencil<T, RC, MK, T1, T2, W1, W2>(a, tmx); });

    const double tol = sizeof(T) == 8 ? 1e-12 : 1e-6;   // the sums run in double for both precisions
    for (long long r = 0; r < N; ++r) {
        if (!(std::fabs(dens[r] - dens_ref[r]) <= tol * (1.0 + std::fabs(dens_ref[r])))) {
            printf("FAIL %s observe: %dx%d %s M=%lld dens[%lld] got %.15g want %.15g\n", name, n1, n2, periodic ? "periodic" : "open", M, r, dens[r], dens_ref[r]);
            return false;
        }
        ++nchecks;
        for (int f = 0; f < NF; ++f) {
            const zc got(G[r * NF + f].x, G[r * NF + f].y), ref = G_ref[r * NF + f];
            if (!(std::abs(got - ref) <= tol * (1.0 + std::abs(ref)))) {
                printf("FAIL %s observe: %dx%d %s M=%lld G[%lld,%d] got (%g,%g) want (%g,%g)\n", name, n1, n2, periodic ? "periodic" : "open", M, r, f, got.real(), got.imag(), ref.real(), ref.imag());
                return false;
            }
            ++nchecks;
        }
    }
    return true;
}

// shapes the library ships (stencil.cu variants 7 / 2, stencil_inst.cuh ObsShape) on small ragged lattices
template <int RC, typename MK>
static bool check_pattern(const char* name) {
    bool ok = true;
    if constexpr (RC == 1) {
        // variant 7: 4x4 tiles, 2x2 warps (8 x 8 cell patches)
        ok = ok && check_apply<double, RC, MK, 4, 4, 2, 2, 1, 3, 1>(name, 11, 9, false, 40, 1, 0);      // open, ragged patches, short last chunk, general values
        ok = ok && check_apply<double, RC, MK, 4, 4, 2, 2, 1, 3, 1>(name, 11, 9, false, 40, 1, 1);      // same, Hermitian values + shared loads
        ok = ok && check_apply<double, RC, MK, 4, 4, 2, 2, 1, 3, 1>(name, 27, 19, true, 96, 2, 3);      // periodic, interior patches through tensor-map boxes
        ok = ok && check_apply<double, RC, MK, 4, 4, 2, 2, 1, 3, 1>(name, 26, 18, false, 40, 1, 3);     // open, boxes + ragged chunk (zero-filled columns)
        ok = ok && check_apply<double, RC, MK, 4, 4, 2, 2, 1, 0, 1>(name, 3, 3, true, 32, 1, 1);        // 3x3 torus: every neighbour is a periodic image
        ok = ok && check_apply<float, RC, MK, 4, 4, 2, 2, 1, 3, 1>(name, 19, 26, true, 136, 1, 3);      // complex64, ragged tail
        ok = ok && check_apply<double, RC, MK, 4, 4, 2, 2, 1, 0, 1>(name, 9, 8, false, 32, 1, 2);       // plain SpMM, general values
        ok = ok && check_apply<double, RC, MK, 4, 4, 1, 2, 1, 3, 1>(name, 13, 19, true, 40, 1, 3);      // 64-thread CTAs (4 x 8 patches)
        ok = ok && check_apply<double, RC, MK, 4, 4, 1, 1, 1, 3, 1>(name, 13, 11, false, 40, 1, 3);     // one warp per CTA
        ok = ok && check_apply<double, RC, MK, 4, 4, 2, 2, 1, 3, 1>(name, 27, 19, true, 96, 2, 7);      // real / imaginary class scalars, shared loads, boxes
        ok = ok && check_apply<double, RC, MK, 4, 4, 2, 2, 1, 3, 1>(name, 11, 9, false, 40, 1, 4);      // class scalars, general body, ragged patches
        ok = ok && check_apply<double, RC, MK, 4, 4, 2, 2, 1, 0, 1>(name, 11, 9, false, 40, 1, 5);      // class scalars, plain SpMM (no self term)
        ok = ok && check_apply<float, RC, MK, 4, 4, 2, 2, 1, 3, 1>(name, 19, 26, true, 136, 1, 7);      // complex64 class scalars
        ok = ok && check_apply<double, RC, MK, 4, 4, 2, 2, 1, 3, 1>(name, 11, 9, true, 40, 1, 11);      // flag cleared: complex values although a class copy exists
        ok = ok && check_observe<double, RC, MK, 2, 2, 4, 2>(name, 11, 9, false, 37, 40, true, 1);
        ok = ok && check_observe<double, RC, MK, 2, 2, 4, 2>(name, 8, 5, true, 130, 136, false, 2);            // pipeline wraps its 3 stages
        ok = ok && check_observe<float, RC, MK, 2, 2, 4, 2>(name, 3, 3, true, 70, 72, true, 4);
        ok = ok && check_observe<double, RC, MK, 2, 2, 4, 2>(name, 27, 14, true, 100, 104, true, 2, true);     // interior patches through tensor-map boxes
        ok = ok && check_observe<float, RC, MK, 2, 2, 4, 2>(name, 26, 13, false, 70, 72, false, 3, true);
    } else if constexpr (RC >= 3) {
        // variant 19: 2x2 tiles, 2x2 warps (4 x 4 cell patches); kagome (3 rows per cell) and spin-1/2 honeycomb (4 rows)
        ok = ok && check_apply<double, RC, MK, 2, 2, 2, 2, 1, 3, 1>(name, 7, 5, false, 40, 1, 0);
        ok = ok && check_apply<double, RC, MK, 2, 2, 2, 2, 1, 3, 1>(name, 7, 5, false, 40, 1, 1);
        ok = ok && check_apply<double, RC, MK, 2, 2, 2, 2, 1, 3, 1>(name, 14, 11, true, 96, 2, 3);
        ok = ok && check_apply<double, RC, MK, 2, 2, 2, 2, 1, 3, 1>(name, 13, 10, false, 40, 1, 3);
        ok = ok && check_apply<double, RC, MK, 2, 2, 2, 2, 1, 0, 1>(name, 3, 3, true, 32, 1, 1);
        ok = ok && check_apply<float, RC, MK, 2, 2, 2, 2, 1, 3, 1>(name, 9, 7, true, 136, 1, 3);
        ok = ok && check_apply<double, RC, MK, 2, 2, 2, 2, 1, 1, 1>(name, 7, 5, true, 32, 1, 3);        // Clenshaw term
        ok = ok && check_apply<double, RC, MK, 2, 2, 2, 2, 1, 2, 1>(name, 7, 5, false, 32, 1, 2);       // Horner term
        ok = ok && check_apply<double, RC, MK, 2, 2, 4, 2, 1, 3, 1>(name, 14, 11, true, 40, 1, 3);      // 256 threads, 8 x 4 patches
        ok = ok && check_apply<double, RC, MK, 2, 2, 2, 2, 1, 3, 1>(name, 14, 11, true, 96, 2, 7);      // real / imaginary class scalars
        ok = ok && check_apply<double, RC, MK, 2, 2, 2, 2, 1, 0, 1>(name, 7, 5, false, 40, 1, 4);
        ok = ok && check_apply<float, RC, MK, 2, 2, 2, 2, 1, 3, 1>(name, 9, 7, true, 136, 1, 7);
        constexpr int NFW = st_nfwd<RC>(MK::mask);
        constexpr int OT2 = (RC == 3 ? NFW > 4 : NFW > 3) ? 1 : 2;
        constexpr int OW2 = (OT2 == 1 && RC == 4) ? 4 : 2;   // four rows, one cell per thread: 4 x 4-cell patches, 512 threads (ObsShape<4, true>)
        ok = ok && check_observe<double, RC, MK, 1, OT2, 4, OW2>(name, 7, 5, false, 37, 40, true, 1);
        ok = ok && check_observe<double, RC, MK, 1, OT2, 4, OW2>(name, 4, 5, true, 130, 136, false, 2);
        ok = ok && check_observe<float, RC, MK, 1, OT2, 4, OW2>(name, 3, 3, true, 70, 72, true, 4);
        ok = ok && check_observe<double, RC, MK, 1, OT2, 4, OW2>(name, 14, 13, true, 100, 104, true, 2, true);
    } else {
        // variant 2: 4x2 tiles, 2x2 warps (8 x 4 cell patches)
        ok = ok && check_apply<double, RC, MK, 4, 2, 2, 2, 1, 3, 1>(name, 11, 5, false, 40, 1, 0);
        ok = ok && check_apply<double, RC, MK, 4, 2, 2, 2, 1, 3, 1>(name, 11, 5, false, 40, 1, 1);
        ok = ok && check_apply<double, RC, MK, 4, 2, 2, 2, 1, 3, 1>(name, 27, 11, true, 96, 2, 3);
        ok = ok && check_apply<double, RC, MK, 4, 2, 2, 2, 1, 3, 1>(name, 26, 10, false, 40, 1, 3);
        ok = ok && check_apply<double, RC, MK, 4, 2, 2, 2, 1, 0, 1>(name, 3, 3, true, 32, 1, 1);
        ok = ok && check_apply<float, RC, MK, 4, 2, 2, 2, 1, 3, 1>(name, 19, 14, true, 136, 1, 3);
        ok = ok && check_apply<double, RC, MK, 4, 2, 2, 2, 1, 0, 1>(name, 9, 4, false, 32, 1, 2);
        ok = ok && check_apply<double, RC, MK, 4, 2, 1, 2, 1, 3, 1>(name, 13, 11, true, 40, 1, 3);      // 64-thread CTAs (4 x 4 patches)
        ok = ok && check_apply<double, RC, MK, 4, 2, 1, 3, 1, 3, 1>(name, 13, 15, false, 40, 1, 3);     // 96 threads, 4 x 6 patches
        ok = ok && check_apply<double, RC, MK, 4, 2, 1, 1, 1, 3, 1>(name, 13, 7, true, 40, 1, 3);       // one warp per CTA
        ok = ok && check_apply<double, RC, MK, 4, 2, 2, 2, 1, 3, 1>(name, 27, 11, true, 96, 2, 7);      // real / imaginary class scalars, shared loads, boxes
        ok = ok && check_apply<double, RC, MK, 4, 2, 2, 2, 1, 3, 1>(name, 11, 5, false, 40, 1, 4);      // class scalars, general body, ragged patches
        ok = ok && check_apply<double, RC, MK, 4, 2, 2, 2, 1, 0, 1>(name, 11, 5, false, 40, 1, 5);      // class scalars, plain SpMM (no self term)
        ok = ok && check_apply<float, RC, MK, 4, 2, 2, 2, 1, 3, 1>(name, 19, 14, true, 136, 1, 7);      // complex64 class scalars
        ok = ok && check_apply<double, RC, MK, 4, 2, 2, 2, 1, 3, 1>(name, 11, 5, true, 40, 1, 11);      // flag cleared: complex values although a class copy exists
        // observables shape of stencil_inst.cuh ObsShape: one cell per thread for wide forward lists
        constexpr int OT2 = st_nfwd<RC>(MK::mask) > 6 ? 1 : 2;
        ok = ok && check_observe<double, RC, MK, 1, OT2, 4, 2>(name, 7, 5, false, 37, 40, true, 1);
        ok = ok && check_observe<double, RC, MK, 1, OT2, 4, 2>(name, 4, 5, true, 130, 136, false, 2);
        ok = ok && check_observe<float, RC, MK, 1, OT2, 4, 2>(name, 3, 3, true, 70, 72, true, 4);
        ok = ok && check_observe<double, RC, MK, 1, OT2, 4, 2>(name, 14, 13, true, 100, 104, true, 2, true);
        ok = ok && check_observe<float, RC, MK, 1, OT2, 4, 2>(name, 13, 11, false, 70, 72, false, 3, true);
    }
    return ok;
}

// LM_EMUL_GROUP (0 .. 3) selects a subset of the patterns at compile time: the fully unrolled kernels are large, so the
// test-suite builds the four groups as separate programs in parallel (undefined = everything in one program)
#ifdef LM_EMUL_GROUP
#define LM_IN_GROUP(g) (LM_EMUL_GROUP == (g))
#else
#define LM_IN_GROUP(g) 1
#endif
int main(int argc, char** argv) {
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    bool ok = true;
#if LM_IN_GROUP(0)
    if (only < 0 || only == 0) ok = ok && check_pattern<1, StPat<0>>("square-nn");
    if (only < 0 || only == 1) ok = ok && check_pattern<1, StPat<1>>("rc1-full");
    if (only < 0 || only == 10) ok = ok && check_pattern<2, StPat<9>>("qwz-diag");
    if (only < 0 || only == 5) {
        // the remaining Clenshaw / Horner modes and the direct-load kernel on one pattern each
        ok = ok && check_apply<double, 2, StPat<4>, 4, 2, 2, 2, 1, 1, 1>("haldane", 13, 9, true, 32, 1, 3);
        ok = ok && check_apply<double, 2, StPat<4>, 4, 2, 2, 2, 1, 2, 1>("haldane", 13, 9, false, 32, 1, 3);
        ok = ok && check_apply<double, 1, StPat<0>, 4, 4, 2, 2, 1, 3, 0>("square-nn", 9, 9, true, 72, 1, 0);
        ok = ok && check_apply<double, 2, StPat<3>, 4, 2, 2, 4, 1, 3, 0>("qwz", 9, 9, false, 40, 1, 0);
    }
#endif
#if LM_IN_GROUP(1)
    if (only < 0 || only == 2) ok = ok && check_pattern<2, StPat<2>>("honeycomb-nn");
    if (only < 0 || only == 3) ok = ok && check_pattern<2, StPat<3>>("qwz");
    if (only < 0 || only == 7) ok = ok && check_pattern<3, StPat<6>>("kagome-nn");
#endif
#if LM_IN_GROUP(2)
    if (only < 0 || only == 4) ok = ok && check_pattern<2, StPat<4>>("haldane");
    if (only < 0 || only == 8) ok = ok && check_pattern<3, StPat<7>>("kagome-nnn");
#endif
#if LM_IN_GROUP(3)
    if (only < 0 || only == 6) ok = ok && check_pattern<2, StPat<5>>("rc2-full");
    if (only < 0 || only == 9) ok = ok && check_pattern<4, StPat<8>>("kanemele");
#endif
    if (!ok) return 1;
    printf("OK %lld checks, %lld bytes through cp.async.bulk\n", nchecks, lm_emul::bulk_bytes());
    return 0;
}
