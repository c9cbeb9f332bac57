/* C-ABI smoke test: a plain C translation unit that includes include/lm_b200.h, links against
 * liblm_b200.so and drives the hot path the way a foreign-language binding (Julia ccall) does -
 * Julia-style 1-based CSC in, column-major Psi in, localdensity / currents out.  Built with gcc by
 * tests/test_c_abi.py: the header is thereby compiled as C (not only as C++ inside api.cu), and every
 * call below is type-checked against its prototypes.  Exit code 0 and "c_abi_smoke OK" on success. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "lm_b200.h"

#define CHECK(call) do { int32_t st_ = (call); if (st_ != LM_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, (int)st_, lm_last_error()); return 1; } } while (0)

int main(void) {
    /* open 6 x 5 square lattice, nearest-neighbour hopping 1 with a Landau-gauge Peierls phase, site (x, y) -> x * 5 + y */
    enum { NX = 6, NY = 5, N = NX * NY, M = 40 };
    const double B = 0.1, two_pi = 6.283185307179586;
    int64_t colptr[N + 1], rowval[4 * N];
    double nzval[2 * 4 * N];
    int64_t nnz = 0;
    for (int j = 0; j < N; ++j) {                     /* column j: rows i ascending (CSC) */
        colptr[j] = nnz + 1;
        const int xj = j / NY, yj = j % NY;
        for (int i = 0; i < N; ++i) {
            const int xi = i / NY, yi = i % NY;
            if (abs(xi - xj) + abs(yi - yj) != 1) continue;
            /* H[i, j] = exp(-2 pi i * (x_i + x_j)(y_j - y_i) B / 2)  (bond i -> j, src/zoo/magneticfields.jl:15) */
            const double phi = -two_pi * 0.5 * ((xi + 1) + (xj + 1)) * (double)(yj - yi) * B;
            rowval[nnz] = i + 1; nzval[2 * nnz] = cos(phi); nzval[2 * nnz + 1] = sin(phi); ++nnz;
        }
    }
    colptr[N] = nnz + 1;

    lm_ctx* ctx = NULL; lm_ham* ham = NULL; lm_state* psi = NULL; lm_state* copy = NULL;
    CHECK(lm_ctx_create(0, LM_C128, NULL, &ctx));
    CHECK(lm_ham_create_csc(ctx, N, 1, colptr, rowval, nzval, 1, &ham));
    CHECK(lm_ham_set_lattice_dims(ham, NX, NY));
    int64_t n_out = 0, nnz_out = 0; int32_t n_int = 0, width = 0;
    CHECK(lm_ham_dims(ham, &n_out, &n_int, &nnz_out, &width));
    if (n_out != N || nnz_out != nnz || n_int != 1) { fprintf(stderr, "lm_ham_dims mismatch\n"); return 1; }
    double emin = 0, emax = 0;
    CHECK(lm_ham_spectral_bounds(ham, &emin, &emax));
    if (!(emin <= -3.9 && emax >= 3.9)) { fprintf(stderr, "Gershgorin bounds [%g, %g]\n", emin, emax); return 1; }

    CHECK(lm_state_create_psi_synth(ctx, N, M, 0, 7u, &psi));
    double n2[M], w_total = 0;
    CHECK(lm_state_column_norms2(psi, n2));
    for (int c = 0; c < M; ++c) w_total += n2[c];
    CHECK(lm_state_copy(psi, &copy));

    int32_t nmv = 0;
    for (int k = 0; k < 5; ++k) CHECK(lm_step(ham, psi, 0.1, 1e-12, LM_METHOD_AUTO, &nmv));
    if (nmv <= 0) { fprintf(stderr, "lm_step reported %d H applications\n", (int)nmv); return 1; }

    int64_t npairs = 0;
    CHECK(lm_currents_npairs(ham, &npairs));
    if (npairs != (NX - 1) * NY + NX * (NY - 1)) { fprintf(stderr, "npairs = %lld\n", (long long)npairs); return 1; }
    int32_t* I = malloc(sizeof(int32_t) * npairs); int32_t* J = malloc(sizeof(int32_t) * npairs);
    double* cur = malloc(sizeof(double) * npairs); double rho[N], rho2[N];
    CHECK(lm_currents_pairs(ham, I, J));
    CHECK(lm_observables(ham, psi, rho, cur));
    CHECK(lm_local_density(psi, 1, rho2));
    double total = 0, dmax = 0;
    for (int i = 0; i < N; ++i) { total += rho[i]; dmax = fmax(dmax, fabs(rho[i] - rho2[i])); }
    if (fabs(total - w_total) > 1e-12 * w_total || dmax > 1e-14) { fprintf(stderr, "trace %g vs %g, lm_local_density deviates by %g\n", total, w_total, dmax); return 1; }
    /* continuity: d rho_i / dt = sum_j J_ij; with unit weights on a unitary evolution sum_ij J_ij over ordered pairs is 0 */
    double one = 0;
    CHECK(lm_bond_currents(ham, psi, 1, &I[0], &J[0], &one));
    if (fabs(one - cur[0]) > 1e-14) { fprintf(stderr, "lm_bond_currents %g vs %g\n", one, cur[0]); return 1; }
    /* stepping back with -dt returns to the copy */
    for (int k = 0; k < 5; ++k) CHECK(lm_step(ham, psi, -0.1, 1e-12, LM_METHOD_AUTO, NULL));
    double* a = malloc(sizeof(double) * 2 * N * M); double* b = malloc(sizeof(double) * 2 * N * M);
    CHECK(lm_state_download_psi(psi, a));
    CHECK(lm_state_download_psi(copy, b));
    double back = 0;
    for (int q = 0; q < 2 * N * M; ++q) back = fmax(back, fabs(a[q] - b[q]));
    if (back > 1e-12) { fprintf(stderr, "round trip deviates by %g\n", back); return 1; }
    int64_t launches = 0;
    CHECK(lm_ctx_launch_count(ctx, &launches));
    CHECK(lm_state_destroy(psi)); CHECK(lm_state_destroy(copy)); CHECK(lm_ham_destroy(ham)); CHECK(lm_ctx_destroy(ctx));
    free(I); free(J); free(cur); free(a); free(b);
    printf("c_abi_smoke OK: version %d, %lld launches, round trip %.1e\n", (int)lm_version(), (long long)launches, back);
    return 0;
}
