// CPU execution of the register-tile body of the stencil kernels (csrc/stencil.cuh: st_tile and the
// compile-time mask helpers) against an independently written reference loop, for every compiled
// pattern and several tile shapes.  Built with plain g++ through the shim in shim/cuda_runtime.h.
// TEST INFRASTRUCTURE ONLY.  Prints "OK <n checks>" or the first mismatch; exit code 0 / 1.
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <random>
#include "stencil.cuh"

using namespace lm;
typedef std::complex<double> zc;
static std::mt19937_64 rng(12345);
static double rnd() { return std::uniform_real_distribution<double>(-1.0, 1.0)(rng); }
static int nchecks = 0;

// independent restatement of the slot order: entries of out row a by (offset o, in row b) ascending
static int ref_slot(int rc, st_mask_t m, int o, int a, int b) {
    int s = 0;
    for (int oo = 0; oo < 9; ++oo)
        for (int bb = 0; bb < rc; ++bb) {
            const bool set = (m.w[(oo * rc * rc + a * rc + bb) >> 6] >> ((oo * rc * rc + a * rc + bb) & 63)) & 1ull;
            if (oo == o && bb == b) return set ? s : -1;
            if (set) ++s;
        }
    return -1;
}

template <typename T, int RC, typename MK, int T1, int T2, bool SELF>
static bool check_tile(const char* name) {
    using E = typename pack<T>::E;
    using T2c = typename cx2<T>::type;
    constexpr int EC = pack<T>::EC;
    constexpr int SW = st_width<RC>(MK::mask);
    static zc x[T1 + 2][T2 + 2][RC][2], h[T1][T2][RC][20];
    for (auto& p : x) for (auto& q : p) for (auto& r : q) for (auto& v : r) v = zc(rnd(), rnd());
    for (auto& p : h) for (auto& q : p) for (auto& r : q) for (auto& v : r) v = zc(rnd(), rnd());
    const zc g(rnd(), rnd());
    E acc[T1][T2][RC][1];
    for (auto& p : acc) for (auto& q : p) for (auto& r : q) pzero(r[0]);
    const T2c gg = cmake<T2c>(g.real(), g.imag());
    st_tile<T, RC, MK, T1, T2, 1, SELF, 0>(acc, gg,
        [&](auto U1, auto U2, auto B, int) {
            constexpr int u1 = decltype(U1)::value, u2 = decltype(U2)::value, b = decltype(B)::value;
            E e;
            if constexpr (EC == 1) { e.x = x[u1][u2][b][0].real(); e.y = x[u1][u2][b][0].imag(); }
            else { e.x = (float)x[u1][u2][b][0].real(); e.y = (float)x[u1][u2][b][0].imag(); e.z = (float)x[u1][u2][b][1].real(); e.w = (float)x[u1][u2][b][1].imag(); }
            return e;
        },
        [&](auto V1, auto V2, auto A, auto S) {
            constexpr int v1 = decltype(V1)::value, v2 = decltype(V2)::value, a = decltype(A)::value, sl = decltype(S)::value;
            static_assert(sl < SW, "slot beyond the stencil width");
            return cmake<T2c>(h[v1][v2][a][sl].real(), h[v1][v2][a][sl].imag());
        });
    const double tol = sizeof(T) == 8 ? 1e-13 : 2e-5;
    for (int v1 = 0; v1 < T1; ++v1) for (int v2 = 0; v2 < T2; ++v2) for (int a = 0; a < RC; ++a) for (int e = 0; e < EC; ++e) {
        zc ref = SELF ? g * x[v1 + 1][v2 + 1][a][e] : zc(0, 0);
        for (int o = 0; o < 9; ++o) for (int b = 0; b < RC; ++b) {
            const int s = ref_slot(RC, MK::mask, o, a, b);
            if (s < 0) continue;
            if (s != st_slot<RC>(MK::mask, o, a, b)) { printf("FAIL %s: st_slot(%d,%d,%d) = %d, reference %d\n", name, o, a, b, st_slot<RC>(MK::mask, o, a, b), s); return false; }
            ref += h[v1][v2][a][s] * x[v1 + 1 + (o / 3 - 1)][v2 + 1 + (o % 3 - 1)][b][e];
        }
        double re, im;
        if constexpr (EC == 1) { re = acc[v1][v2][a][0].x; im = acc[v1][v2][a][0].y; }
        else { re = e ? acc[v1][v2][a][0].z : acc[v1][v2][a][0].x; im = e ? acc[v1][v2][a][0].w : acc[v1][v2][a][0].y; }
        if (std::abs(zc(re, im) - ref) > tol * (1.0 + std::abs(ref))) {
            printf("FAIL %s: tile %dx%d self=%d out (%d,%d,%d) col %d: got (%g,%g) want (%g,%g)\n", name, T1, T2, (int)SELF, v1, v2, a, e, re, im, ref.real(), ref.imag());
            return false;
        }
        ++nchecks;
    }
    return true;
}

template <int RC, typename MK>
static bool check_helpers(const char* name) {
    // width = widest row; forward entries = one per bond direction towards a later cell / later row
    int w = 1, nf = 1;
    for (int a = 0; a < RC; ++a) {
        int s = 0, f = 0;
        for (int o = 0; o < 9; ++o) for (int b = 0; b < RC; ++b) {
            const bool set = st_bit<RC>(MK::mask, o, a, b);
            if (!set) continue;
            ++s;
            if (o > 4 || (o == 4 && b > a)) {
                if (st_fslot<RC>(MK::mask, o, a, b) != f) { printf("FAIL %s: st_fslot\n", name); return false; }
                ++f;
            }
            // Hermitian pattern: the reverse bond exists
            const bool rev = st_bit<RC>(MK::mask, 8 - o, b, a);
            if (!rev) { printf("FAIL %s: pattern not symmetric at (%d,%d,%d)\n", name, o, a, b); return false; }
        }
        if (s > w) w = s;
        if (f > nf) nf = f;
    }
    if (w != st_width<RC>(MK::mask) || nf != st_nfwd<RC>(MK::mask)) { printf("FAIL %s: width / forward count\n", name); return false; }
    ++nchecks;
    return true;
}

template <int RC, typename MK>
static bool check_pattern(const char* name) {
    bool ok = check_helpers<RC, MK>(name);
    if constexpr (RC >= 3)        // three / four rows per cell run 2 x 2 tiles (stencil.cu variant 19)
        return ok && check_tile<double, RC, MK, 2, 2, true>(name) && check_tile<double, RC, MK, 1, 2, false>(name) && check_tile<float, RC, MK, 2, 2, true>(name);
    ok = ok && check_tile<double, RC, MK, 4, 2, true>(name) && check_tile<double, RC, MK, 2, 2, false>(name);
    ok = ok && check_tile<double, RC, MK, 1, 2, true>(name) && check_tile<float, RC, MK, 4, 2, true>(name);
    if (RC == 1) ok = ok && check_tile<double, RC, MK, 4, 4, true>(name);
    return ok;
}

int main() {
    // LM_EMUL_GROUP (0 .. 2): the test-suite builds the pattern groups as separate programs in parallel
    bool ok = true;
#if !defined(LM_EMUL_GROUP) || LM_EMUL_GROUP == 0
    ok = ok && check_pattern<1, StPat<0>>("square-nn");
    ok = ok && check_pattern<1, StPat<1>>("rc1-full");
    ok = ok && check_pattern<2, StPat<2>>("honeycomb-nn");
    ok = ok && check_pattern<2, StPat<3>>("qwz");
    ok = ok && check_pattern<2, StPat<9>>("qwz-diag");
#endif
#if !defined(LM_EMUL_GROUP) || LM_EMUL_GROUP == 1
    ok = ok && check_pattern<2, StPat<4>>("haldane");
    ok = ok && check_pattern<2, StPat<5>>("rc2-full");
#endif
#if !defined(LM_EMUL_GROUP) || LM_EMUL_GROUP == 2
    ok = ok && check_pattern<3, StPat<6>>("kagome-nn");
    ok = ok && check_pattern<3, StPat<7>>("kagome-nnn");
    ok = ok && check_pattern<4, StPat<8>>("kanemele");
#endif
    if (!ok) return 1;
    printf("OK %d checks\n", nchecks);
    return 0;
}
