// svd_host.h -- the one piece of dense linear algebra the lossy B strategy needs on the host: the
// largest singular value of the interferometer (reference: LAPACKE_zgesvd in
// CGeneralizedCliffordsBLossySimulationStrategy.cpp:386-452, of which only max(s) and U S V^H / max(s)
// = M / max(s) are used).  LAPACK is not a dependency of this library.
#pragma once
#include <math.h>
#include <stddef.h>

#include <algorithm>
#include <vector>

namespace pqb200 {

// Largest eigenvalue of the Hermitian matrix H = M^H M (n x n) by cyclic complex Jacobi rotations:
// the squared largest singular value of M.  Host-side, O(n^3) per sweep, n <= 255.
inline double largest_singular_value_sq(const double* a, size_t n, size_t stride) {
    std::vector<double> hr(n * n, 0.0), hi(n * n, 0.0);
    for (size_t i = 0; i < n; i++)
        for (size_t j = 0; j < n; j++) {
            long double sr = 0.0L, si = 0.0L;
            for (size_t k = 0; k < n; k++) {
                const long double ar = a[2 * (k * stride + i)], ai = a[2 * (k * stride + i) + 1];
                const long double br = a[2 * (k * stride + j)], bi = a[2 * (k * stride + j) + 1];
                sr += ar * br + ai * bi;  // conj(a_ki) * a_kj
                si += ar * bi - ai * br;
            }
            hr[i * n + j] = (double)sr;
            hi[i * n + j] = (double)si;
        }
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0;
        for (size_t p = 0; p < n; p++)
            for (size_t q = p + 1; q < n; q++) off += hr[p * n + q] * hr[p * n + q] + hi[p * n + q] * hi[p * n + q];
        double diag = 0.0;
        for (size_t p = 0; p < n; p++) diag += hr[p * n + p] * hr[p * n + p];
        if (off <= 1e-32 * diag) break;
        for (size_t p = 0; p < n; p++)
            for (size_t q = p + 1; q < n; q++) {
                const double xr = hr[p * n + q], xi = hi[p * n + q];
                const double ab = sqrt(xr * xr + xi * xi);
                if (ab < 1e-300) continue;
                // H_pq = ab * e^{i phi}; rotate with c, s e^{i phi}
                const double er = xr / ab, ei = xi / ab;
                const double app = hr[p * n + p], aqq = hr[q * n + q];
                const double tau = (aqq - app) / (2.0 * ab);
                const double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                const double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
                // columns: H <- H J, J = [[c, s e^{i phi}], [-s e^{-i phi}, c]] on (p, q)
                for (size_t k = 0; k < n; k++) {
                    const double kpr = hr[k * n + p], kpi = hi[k * n + p];
                    const double kqr = hr[k * n + q], kqi = hi[k * n + q];
                    // new_p = c*kp - s*e^{-i phi}*kq ; new_q = s*e^{i phi}*kp + c*kq
                    hr[k * n + p] = c * kpr - s * (er * kqr + ei * kqi);
                    hi[k * n + p] = c * kpi - s * (er * kqi - ei * kqr);
                    hr[k * n + q] = s * (er * kpr - ei * kpi) + c * kqr;
                    hi[k * n + q] = s * (er * kpi + ei * kpr) + c * kqi;
                }
                // rows: H <- J^H H
                for (size_t k = 0; k < n; k++) {
                    const double pkr = hr[p * n + k], pki = hi[p * n + k];
                    const double qkr = hr[q * n + k], qki = hi[q * n + k];
                    // new_p = c*pk - s*e^{i phi}*qk ; new_q = s*e^{-i phi}*pk + c*qk
                    hr[p * n + k] = c * pkr - s * (er * qkr - ei * qki);
                    hi[p * n + k] = c * pki - s * (er * qki + ei * qkr);
                    hr[q * n + k] = s * (er * pkr + ei * pki) + c * qkr;
                    hi[q * n + k] = s * (er * pki - ei * pkr) + c * qki;
                }
            }
    }
    double mx = 0.0;
    for (size_t i = 0; i < n; i++) mx = std::max(mx, hr[i * n + i]);
    return mx;
}


}  // namespace pqb200
