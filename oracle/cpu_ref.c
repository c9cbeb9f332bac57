/*
 * oracle/cpu_ref.c - C restatement of the reference's CPU evolution of a block of kets.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py): never linked into, imported
 * or executed by the product path.  Used by tests/ (cross-checked against the numpy oracle)
 * and by bench.py's cpu_baseline / `--impl reference` legs as the timed CPU baseline.
 *
 * What it restates (paths relative to /root/reference):
 *   Evolution(KrylovKitExp(), H, kets...) evolves every registered Ket separately
 *   (src/evolution.jl:245-247) with KrylovKit.exponentiate(H, -im*dt, ket)
 *   (src/evolution.jl:150-154).  KrylovKit.jl is not vendored in the reference tree
 *   (Project.toml:10,26, compat "0.4 - 0.9", no Manifest): its published algorithm - Lanczos
 *   exponential integrator, defaults krylovdim = 30, tol = 1e-12, full re-orthogonalisation
 *   (orth = ModifiedGramSchmidtIR), residual estimate beta_m |[exp(tau T_m)]_{m,1}| - is
 *   restated here exactly as in oracle/evolution.py::krylov_exponentiate.
 *   localdensity (src/operators/latticeutils.jl:41-45) for the per-frame density.
 * The reference itself is single-threaded Julia; columns are independent, so the pthread loop
 * over columns is the most favourable CPU arrangement (thread count is reported).
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

typedef double complex zc;

static void spmv(int64_t n, const int64_t* rowptr, const int32_t* col, const zc* val, const zc* x, zc* y) {
    for (int64_t i = 0; i < n; ++i) {
        zc s = 0;
        for (int64_t q = rowptr[i]; q < rowptr[i + 1]; ++q) s += val[q] * x[col[q]];
        y[i] = s;
    }
}

/* symmetric tridiagonal eigen-decomposition, implicit QL (EISPACK tql2): d diag (in: diag,
 * out: eigenvalues), e off-diag (e[0..m-2]), z m x m row-major eigenvectors (in: identity) */
static int tql2(int m, double* d, double* e, double* z) {
    if (m == 1) return 0;
    e[m - 1] = 0.0;
    for (int l = 0; l < m; ++l) {
        int iter = 0, mm;
        do {
            for (mm = l; mm < m - 1; ++mm) {
                double dd = fabs(d[mm]) + fabs(d[mm + 1]);
                if (fabs(e[mm]) <= 2.3e-16 * dd) break;
            }
            if (mm != l) {
                if (iter++ == 200) return -1;
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = hypot(g, 1.0);
                g = d[mm] - d[l] + e[l] / (g + (g >= 0 ? fabs(r) : -fabs(r)));
                double s = 1.0, c = 1.0, p = 0.0;
                int i;
                for (i = mm - 1; i >= l; --i) {
                    double f = s * e[i], b = c * e[i];
                    e[i + 1] = (r = hypot(f, g));
                    if (r == 0.0) { d[i + 1] -= p; e[mm] = 0.0; break; }
                    s = f / r; c = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    d[i + 1] = g + (p = s * r);
                    g = c * r - b;
                    for (int k = 0; k < m; ++k) {
                        f = z[k * m + i + 1];
                        z[k * m + i + 1] = s * z[k * m + i] + c * f;
                        z[k * m + i] = c * z[k * m + i] - s * f;
                    }
                }
                if (r == 0.0 && i >= l) continue;
                d[l] -= p; e[l] = g; e[mm] = 0.0;
            }
        } while (mm != l);
    }
    return 0;
}

/* y = exp(tau * T) e_1 for T = tridiag(alpha, beta), tau purely imaginary = -i*dt */
static void tridiag_exp_col(int m, const double* alpha, const double* beta, double dt, zc* y) {
    double d[64], e[64], z[64 * 64];
    for (int i = 0; i < m; ++i) { d[i] = alpha[i]; e[i] = (i < m - 1) ? beta[i] : 0.0; }
    memset(z, 0, sizeof(double) * m * m);
    for (int i = 0; i < m; ++i) z[i * m + i] = 1.0;
    tql2(m, d, e, z);
    for (int i = 0; i < m; ++i) {
        zc s = 0;
        for (int k = 0; k < m; ++k) s += z[i * m + k] * cexp(-I * dt * d[k]) * z[0 * m + k];
        y[i] = s;
    }
}

/* exponentiate one column in place; returns number of matvecs, negative if not converged */
static int64_t krylov_column(int64_t n, const int64_t* rowptr, const int32_t* col, const zc* val,
                             zc* v, double dt_total, int krylovdim, double tol, int maxiter, zc* work) {
    zc* V = work;                       /* (krylovdim) x n basis */
    zc* r = work + (size_t)krylovdim * n;
    double alpha[64], beta[64];
    zc y[64];
    double done = 0.0;
    int64_t nmv = 0;
    if (krylovdim > 60) krylovdim = 60;
    for (int it = 0; it < maxiter; ++it) {
        double remaining = dt_total - done;
        if (remaining == 0.0) return nmv;
        double beta0 = 0.0;
        for (int64_t i = 0; i < n; ++i) beta0 += creal(v[i]) * creal(v[i]) + cimag(v[i]) * cimag(v[i]);
        beta0 = sqrt(beta0);
        if (beta0 == 0.0) return nmv;
        for (int64_t i = 0; i < n; ++i) V[i] = v[i] / beta0;
        int m, converged = 0;
        double b = 0.0;
        for (m = 1; m <= krylovdim; ++m) {
            zc* vm = V + (size_t)(m - 1) * n;
            spmv(n, rowptr, col, val, vm, r);
            ++nmv;
            double a = 0.0;
            for (int64_t i = 0; i < n; ++i) a += creal(conj(vm[i]) * r[i]);
            for (int64_t i = 0; i < n; ++i) r[i] -= a * vm[i];
            if (m > 1) { zc* vp = V + (size_t)(m - 2) * n; for (int64_t i = 0; i < n; ++i) r[i] -= beta[m - 2] * vp[i]; }
            for (int q = 0; q < m; ++q) {           /* full re-orthogonalisation */
                zc* vq = V + (size_t)q * n; zc s = 0;
                for (int64_t i = 0; i < n; ++i) s += conj(vq[i]) * r[i];
                for (int64_t i = 0; i < n; ++i) r[i] -= s * vq[i];
            }
            b = 0.0;
            for (int64_t i = 0; i < n; ++i) b += creal(r[i]) * creal(r[i]) + cimag(r[i]) * cimag(r[i]);
            b = sqrt(b);
            alpha[m - 1] = a;
            tridiag_exp_col(m, alpha, beta, remaining, y);
            double err = fabs(b) * cabs(y[m - 1]) * beta0;
            if (err <= tol * fabs(remaining) / fabs(dt_total) || b < 1e-300) { converged = 1; break; }
            if (m < krylovdim) { beta[m - 1] = b; zc* vn = V + (size_t)m * n; for (int64_t i = 0; i < n; ++i) vn[i] = r[i] / b; }
        }
        if (!converged) {
            m = krylovdim;
            double tau = remaining;
            for (int h = 0; h < 60; ++h) {
                tau *= 0.5;
                tridiag_exp_col(m, alpha, beta, tau, y);
                if (fabs(b) * cabs(y[m - 1]) * beta0 <= tol * fabs(tau) / fabs(dt_total)) break;
            }
            done += tau;
        }
        for (int64_t i = 0; i < n; ++i) {
            zc s = 0;
            for (int q = 0; q < m; ++q) s += V[(size_t)q * n + i] * y[q];
            v[i] = beta0 * s;
        }
        if (converged) return nmv;
    }
    return -nmv;
}

/* Evolve M kets (columns of psi, column-major n x M) by exp(-i H dt); H in CSR (0-based).
 * Columns are distributed over `nthreads` pthreads through an atomic counter.
 * Returns total matvecs (negative if any column failed to converge). */
typedef struct {
    int64_t n, M; const int64_t* rowptr; const int32_t* col; const zc* val; zc* psi;
    double dt, tol; int krylovdim, maxiter;
    int64_t next; int64_t total; int failed; pthread_mutex_t mu;
} job_t;

static void* worker(void* arg) {
    job_t* j = (job_t*)arg;
    zc* work = (zc*)malloc(sizeof(zc) * (size_t)(j->krylovdim + 1) * j->n);
    int64_t mine = 0; int bad = 0;
    for (;;) {
        int64_t c = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (c >= j->M) break;
        int64_t k = krylov_column(j->n, j->rowptr, j->col, j->val, j->psi + (size_t)c * j->n, j->dt,
                                  j->krylovdim, j->tol, j->maxiter, work);
        if (k < 0) { bad = 1; k = -k; }
        mine += k;
    }
    free(work);
    pthread_mutex_lock(&j->mu); j->total += mine; j->failed |= bad; pthread_mutex_unlock(&j->mu);
    return NULL;
}

int lm_ref_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

int64_t lm_ref_krylov_block_step(int64_t n, const int64_t* rowptr, const int32_t* col, const double* val_ri,
                                 int64_t M, double* psi_ri, double dt, int krylovdim, double tol, int maxiter,
                                 int nthreads) {
    job_t j;
    j.n = n; j.M = M; j.rowptr = rowptr; j.col = col; j.val = (const zc*)val_ri; j.psi = (zc*)psi_ri;
    j.dt = dt; j.tol = tol; j.krylovdim = krylovdim; j.maxiter = maxiter;
    j.next = 0; j.total = 0; j.failed = 0;
    pthread_mutex_init(&j.mu, NULL);
    if (nthreads <= 0) nthreads = lm_ref_max_threads();
    if (nthreads > M) nthreads = (int)(M > 0 ? M : 1);
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], NULL, worker, &j);
    worker(&j);
    for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);
    pthread_mutex_destroy(&j.mu);
    return j.failed ? -j.total : j.total;
}

/* rho_i = sum_alpha sum_c w_c |psi[i*n_int + alpha, c]|^2  (src/operators/latticeutils.jl:41-45) */
void lm_ref_localdensity(int64_t n, int64_t M, int n_int, const double* psi_ri, const double* w, double* rho) {
    const zc* psi = (const zc*)psi_ri;
    int64_t ns = n / n_int;
    for (int64_t s = 0; s < ns; ++s) rho[s] = 0.0;
    for (int64_t c = 0; c < M; ++c) {
        double wc = w ? w[c] : 1.0;
        for (int64_t i = 0; i < n; ++i) {
            zc a = psi[(size_t)c * n + i];
            rho[i / n_int] += wc * (creal(a) * creal(a) + cimag(a) * cimag(a));
        }
    }
}

