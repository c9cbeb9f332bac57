// Host-side dense linear algebra for the SMALL matrices of the device eigensolver (lm_eigs_lowest):
// Hermitian eigen-decomposition of an n x n complex matrix, n <= a few hundred (cyclic Jacobi).
// The N x n blocks never leave the device; only n x n Gram / Rayleigh-Ritz matrices come here.
#pragma once
#include <algorithm>
#include <cmath>
#include <complex>
#include <numeric>
#include <vector>

namespace lm {

// A (n x n, row-major, Hermitian; only read) = V diag(w) V^H, eigenvalues ascending, V column j = eigenvector j
// (V row-major: V[i * n + j]).  Returns the number of sweeps used, -1 if not converged.
inline int heig_jacobi(int n, const std::vector<std::complex<double>>& A_in, std::vector<double>& w, std::vector<std::complex<double>>& V) {
    typedef std::complex<double> zc;
    std::vector<zc> A(A_in);
    V.assign((size_t)n * n, zc(0, 0));
    for (int i = 0; i < n; ++i) V[(size_t)i * n + i] = zc(1, 0);
    // symmetrise the input (Gram matrices arrive Hermitian up to rounding)
    for (int i = 0; i < n; ++i) {
        A[(size_t)i * n + i] = zc(A[(size_t)i * n + i].real(), 0.0);
        for (int j = i + 1; j < n; ++j) {
            const zc m = 0.5 * (A[(size_t)i * n + j] + std::conj(A[(size_t)j * n + i]));
            A[(size_t)i * n + j] = m; A[(size_t)j * n + i] = std::conj(m);
        }
    }
    double scale = 0.0;
    for (const zc& v : A) scale = std::max(scale, std::abs(v));
    if (scale == 0.0) { w.assign(n, 0.0); return 0; }
    int sweep = 0;
    for (; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < n; ++p) for (int q = p + 1; q < n; ++q) off = std::max(off, std::abs(A[(size_t)p * n + q]));
        if (off <= 1e-15 * scale) break;
        for (int p = 0; p < n; ++p)
            for (int q = p + 1; q < n; ++q) {
                const zc apq = A[(size_t)p * n + q];
                const double r = std::abs(apq);
                if (r <= 1e-300 || r <= 1e-17 * scale) continue;
                const zc ph = apq / r;                                   // e^{i phi}
                const double tau = (A[(size_t)q * n + q].real() - A[(size_t)p * n + p].real()) / (2.0 * r);
                const double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
                const double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
                // R = [[c, s], [-s e^{-i phi}, c e^{-i phi}]] on (p, q):  A <- R^H A R,  V <- V R
                const zc rqp = -s * std::conj(ph), rqq = c * std::conj(ph);
                for (int k = 0; k < n; ++k) {                            // columns p, q of A and V
                    const zc akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
                    A[(size_t)k * n + p] = akp * c + akq * rqp;
                    A[(size_t)k * n + q] = akp * s + akq * rqq;
                    const zc vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
                    V[(size_t)k * n + p] = vkp * c + vkq * rqp;
                    V[(size_t)k * n + q] = vkp * s + vkq * rqq;
                }
                for (int k = 0; k < n; ++k) {                            // rows p, q of A
                    const zc apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
                    A[(size_t)p * n + k] = c * apk + std::conj(rqp) * aqk;
                    A[(size_t)q * n + k] = s * apk + std::conj(rqq) * aqk;
                }
                A[(size_t)p * n + q] = zc(0, 0); A[(size_t)q * n + p] = zc(0, 0);
                A[(size_t)p * n + p] = zc(A[(size_t)p * n + p].real(), 0.0);
                A[(size_t)q * n + q] = zc(A[(size_t)q * n + q].real(), 0.0);
            }
    }
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](int a, int b) { return A[(size_t)a * n + a].real() < A[(size_t)b * n + b].real(); });
    w.resize(n);
    std::vector<zc> Vs((size_t)n * n);
    for (int j = 0; j < n; ++j) {
        w[j] = A[(size_t)order[j] * n + order[j]].real();
        for (int i = 0; i < n; ++i) Vs[(size_t)i * n + j] = V[(size_t)i * n + order[j]];
    }
    V.swap(Vs);
    return sweep < 60 ? sweep : -1;
}

}  // namespace lm
