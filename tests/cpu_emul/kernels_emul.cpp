// CPU execution of the non-stencil device code of csrc/kernels.cuh (Peierls phase regeneration,
// value assembly, the generic ELL propagator term, the fused localdensity / bond-correlator
// reductions with their finalisation, the Gershgorin enclosure, the per-column Lanczos
// exponential coefficients) behind a plain C interface, so that tests/test_device_code_cpu.py can
// feed them numpy arrays and compare with the oracle - on the CPU, no GPU needed.  The kernels are
// the product's own source, compiled by g++ through shim/cuda_runtime.h (threads of a CTA are
// fibers).  TEST INFRASTRUCTURE ONLY - never linked into the product.
#include "kernels.cuh"

using namespace lm;

namespace {
struct PhaseArgs { long long nb; const double* r; const double2* bfac; int nfields; const int* kinds; const double* params; double2* phase; };
void run_phase(const PhaseArgs a) { k_bond_phase(a.nb, a.r, a.bfac, a.nfields, a.kinds, a.params, a.phase); }

struct AsmArgs { long long E; const double2* stat; const int* cptr; const int* cbond; const double2* camp; const double2* phase; double2* vals; };
void run_assemble(const AsmArgs a) { k_assemble<double>(a.E, a.stat, a.cptr, a.cbond, a.camp, a.phase, a.vals); }

struct GershArgs { long long N; int W; const int* cols; const double2* vals; double* partial; };
void run_gersh(const GershArgs a) { k_gershgorin<double>(a.N, a.W, a.cols, a.vals, a.partial); }

struct ObsArgs { long long N, M, ld; const double2* x; const double* w; const int* cols; int W; const unsigned char* upper; int k0, write_dens; double* dens; double2* G; };
template <int WB> void run_observe(const ObsArgs a) { k_observe<double, WB, false>(a.N, a.M, a.ld, a.x, a.w, a.cols, a.W, a.upper, a.k0, a.write_dens, a.dens, a.G); }
template <int WB> void run_observe_team(const ObsArgs a) { k_observe<double, WB, true>(a.N, a.M, a.ld, a.x, a.w, a.cols, a.W, a.upper, a.k0, a.write_dens, a.dens, a.G); }

struct FinArgs { long long n_sites; int n_int; const double* dens; long long npairs; const int* pair_ptr; const int* pair_ent; const double2* vals; const double2* G; double* obs; };
void run_finalize(const FinArgs a) { k_finalize_obs<double>(a.n_sites, a.n_int, a.dens, a.npairs, a.pair_ptr, a.pair_ent, a.vals, a.G, a.obs, 1, 1); }

struct CoefArgs { long long M; int m; long long ldc; double dt; const double2* alpha; const double* beta; double2* coef; double* err; };
void run_coef(const CoefArgs a) { k_lanczos_coef(a.M, a.m, a.ldc, a.dt, a.alpha, a.beta, a.coef, a.err); }
struct MaxArgs { long long M; const double* err; unsigned long long* out; };
void run_max(const MaxArgs a) { k_max_cols(a.M, a.err, a.out); }
}  // namespace

extern "C" {

// phase[b] = bfac[b] * exp(-2 pi i sum_f line_integral_f(r1_b, r2_b))
void emul_bond_phase(long long nb, const double* r, const double* bfac, int nfields, const int* kinds, const double* params, double* phase) {
    const unsigned th = 64;
    lm_emul::launch(run_phase, dim3((unsigned)((nb + th - 1) / th)), th, PhaseArgs{nb, r, (const double2*)bfac, nfields, kinds, params, (double2*)phase});
}
// vals[e] = static[e] + sum_c amp_c * (conj?) phase[bond_c]
void emul_assemble(long long E, const double* stat, const int* cptr, const int* cbond, const double* camp, const double* phase, double* vals) {
    const unsigned th = 64;
    lm_emul::launch(run_assemble, dim3((unsigned)((E + th - 1) / th)), th, AsmArgs{E, (const double2*)stat, cptr, cbond, (const double2*)camp, (const double2*)phase, (double2*)vals});
}
// folded Gershgorin enclosure (emin, emax, norm_inf) and Hermiticity defect max |H_ij - conj(H_ji)| exactly as
// the host folds the per-CTA partials (4 doubles per CTA)
void emul_gershgorin(long long N, int W, const int* cols, const double* vals, double* out4) {
    const unsigned grid = (unsigned)((N + 255) / 256);
    std::vector<double> partial(4 * (size_t)grid);
    lm_emul::launch(run_gersh, dim3(grid), 256, GershArgs{N, W, cols, (const double2*)vals, partial.data()});
    double lo = 1e300, hi = -1e300, nr = 0.0, as = 0.0;
    for (unsigned g = 0; g < grid; ++g) { lo = std::fmin(lo, partial[4 * g]); hi = std::fmax(hi, partial[4 * g + 1]); nr = std::fmax(nr, partial[4 * g + 2]); as = std::fmax(as, partial[4 * g + 3]); }
    out4[0] = lo; out4[1] = hi; out4[2] = nr; out4[3] = as;
}
// y = alpha H x + gamma x + beta z + delta u through k_apply (generic ELL path, complex128) with
// the launch geometry of api.cu apply(); mode as in the library (0 plain, 1 +beta z, 2 general, 3 factor)
int emul_apply_ell(long long N, int W, const int* cols, const double* vals, long long ld, const double* x, double* y,
                   const double* z, const double* u, const double* abgd /* alpha, gamma, beta, delta (re, im) */, int mode, int cpt, unsigned tps_req) {
    ApplyArgs a;
    a.cols = cols; a.vals = vals; a.W = W; a.N = N; a.ld = ld;
    a.x = x; a.y = y; a.z = z; a.u = u;
    a.pdl = 0;
    a.alpha[0] = abgd[0]; a.alpha[1] = abgd[1]; a.gamma[0] = abgd[2]; a.gamma[1] = abgd[3];
    a.beta[0] = abgd[4]; a.beta[1] = abgd[5]; a.delta[0] = abgd[6]; a.delta[1] = abgd[7];
    int lc = 0; while ((1LL << lc) < ld && lc < 5) lc++;
    const int LC = 1 << lc, LR = 32 >> lc;
    a.lc_log2 = lc;
    const long long tiles_r = (N + 8LL * LR - 1) / (8LL * LR);
    const long long tiles_c = (ld + (long long)LC * cpt - 1) / ((long long)LC * cpt);
    long long tps = std::max<long long>(1, std::min<long long>(tps_req, tiles_c));
    const long long strips = (tiles_c + tps - 1) / tps;
    tps = (tiles_c + strips - 1) / strips;
    a.tiles_c = (unsigned)tiles_c; a.tps = (unsigned)tps;
    dim3 grid((unsigned)(tiles_r * tps), (unsigned)strips);
#define LM_EMUL_APPLY(C, MD) if (cpt == C && mode == MD) { lm_emul::launch(k_apply<double, C, 0, MD>, grid, 256, a); return 0; }
    LM_EMUL_APPLY(1, 0) LM_EMUL_APPLY(1, 1) LM_EMUL_APPLY(1, 2) LM_EMUL_APPLY(1, 3)
    LM_EMUL_APPLY(2, 0) LM_EMUL_APPLY(2, 3) LM_EMUL_APPLY(4, 2) LM_EMUL_APPLY(4, 3)
#undef LM_EMUL_APPLY
    return -1;
}
// fused reductions of api.cu observe() (ELL path) + k_finalize_obs: obs = [rho (n_sites) | J (npairs)]
void emul_observables(long long N, long long M, long long ld, const double* x, const double* w, const int* cols, int W, const unsigned char* upper,
                      const double* vals, long long n_sites, int n_int, long long npairs, const int* pair_ptr, const int* pair_ent,
                      int cta_team, double* dens, double* G, double* obs) {
    ObsArgs a{N, M, ld, (const double2*)x, w, cols, W, upper, 0, 1, dens, (double2*)G};
    const dim3 grid(cta_team ? (unsigned)N : (unsigned)((N + 7) / 8));
    auto go = [&](auto wb) {
        constexpr int WB = decltype(wb)::value;
        if (cta_team) lm_emul::launch(run_observe_team<WB>, grid, 256, a); else lm_emul::launch(run_observe<WB>, grid, 256, a);
    };
    if (W <= 4) go(std::integral_constant<int, 4>{});
    else if (W <= 8) go(std::integral_constant<int, 8>{});
    else for (int k0 = 0; k0 < W; k0 += 16) { a.k0 = k0; a.write_dens = (k0 == 0); go(std::integral_constant<int, 16>{}); }
    const long long tot = n_sites + npairs;
    lm_emul::launch(run_finalize, dim3((unsigned)((tot + 255) / 256)), 256,
                    FinArgs{n_sites, n_int, dens, npairs, pair_ptr, pair_ent, (const double2*)vals, (const double2*)G, obs});
}
// per-column Lanczos exponential coefficients + the convergence reduction (returns max_c err[c])
double emul_lanczos_coef(long long M, int m, long long ldc, double dt, const double* alpha, const double* beta, double* coef, double* err) {
    const unsigned th = 64, grid = (unsigned)((M + th - 1) / th);
    lm_emul::launch(run_coef, dim3(grid), th, CoefArgs{M, m, ldc, dt, (const double2*)alpha, beta, (double2*)coef, err});
    unsigned long long bits = 0;
    lm_emul::launch(run_max, dim3(grid), th, MaxArgs{M, err, &bits});
    double v; std::memcpy(&v, &bits, sizeof v);
    return v;
}

// Fused finalize + peer-memory all-gather of the per-frame [rho | J] (k_finalize_obs_p2p /
// k_obs_p2p_reduce), `nranks` ranks played in ONE address space: every rank owns a symmetric buffer
// flags[2][nranks] | slots[2][nranks][cap] laid out as api.cu does (4096 flag bytes), pushes its
// partial into its slot on EVERY peer and publishes the frame epoch; then every rank acquires all
// flags and sums.  Inputs per rank r: dens_r [N], G_r [N*W]; output per rank: obs_r [tot].
// Returns the number of ranks whose `done` counter was re-armed (must be nranks).
int emul_p2p_frame(int nranks, long long cap, unsigned long long epoch, long long n_sites, int n_int, long long npairs,
                   const int* pair_ptr, const int* pair_ent, const double* vals, const double* const* dens, const double* const* G,
                   void* const* bufs /* nranks symmetric buffers */, unsigned* const* done, double* const* obs) {
    const size_t flag_bytes = 4096;
    const int parity = (int)(epoch & 1);
    const long long tot = n_sites + npairs;
    const int th = 256; const unsigned grid = (unsigned)((tot + th - 1) / th);
    PeerPtrs pp;
    for (int r = 0; r < 8; ++r) {
        char* base = (char*)bufs[r < nranks ? r : 0];
        pp.flags[r] = (unsigned long long*)base; pp.slots[r] = (double*)(base + flag_bytes);
    }
    for (int r = 0; r < nranks; ++r)
        lm_emul::enqueue(grid, th, [](auto... a_) { k_finalize_obs_p2p<double>(a_...); }, n_sites, n_int, dens[r], npairs, pair_ptr, pair_ent,
                         (const double2*)vals, (const double2*)G[r], 1, pp, r, nranks, cap, parity, epoch, done[r]);
    int rearmed = 0;
    for (int r = 0; r < nranks; ++r) {
        rearmed += (*done[r] == 0);
        lm_emul::enqueue(grid, th, [](auto... a_) { k_obs_p2p_reduce(a_...); }, tot, (const double*)((char*)bufs[r] + flag_bytes),
                         (const unsigned long long*)bufs[r], nranks, cap, parity, epoch, obs[r]);
    }
    return rearmed;
}

}  // extern "C"
