// oracle/ref_driver/blas_shim.cpp  --  TEST INFRASTRUCTURE ONLY (not part of the product).
//
// The reference's lossy B strategy calls two BLAS/LAPACK routines that are not installed in this
// image: cblas_zgemm (through common/source/dot.cpp) and LAPACKE_zgesvd
// (CGeneralizedCliffordsBLossySimulationStrategy.cpp:40-47,386-452).  This file provides plain
// textbook implementations with the signatures the reference declares, so that the reference's
// strategy sources compile and run UNCHANGED inside oracle/_ref/libpqbref.so.  Written for this
// repository; nothing here is reference code.
#include <cmath>
#include <complex>
#include <cstring>
#include <vector>

typedef std::complex<double> cplx;

extern "C" {

enum { kRowMajor = 101, kNoTrans = 111, kTrans = 112, kConjTrans = 113 };

// C = alpha * op(A) * op(B) + beta * C, row-major only (the only layout dot.cpp uses)
void cblas_zgemm(int order, int transA, int transB, int M, int N, int K, const void* alpha_,
                 const void* A_, int lda, const void* B_, int ldb, const void* beta_, void* C_, int ldc) {
    const cplx alpha = *(const cplx*)alpha_, beta = *(const cplx*)beta_;
    const cplx* A = (const cplx*)A_;
    const cplx* B = (const cplx*)B_;
    cplx* C = (cplx*)C_;
    (void)order;
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++) {
            cplx acc = 0.0;
            for (int k = 0; k < K; k++) {
                cplx a = transA == kNoTrans ? A[(size_t)i * lda + k] : A[(size_t)k * lda + i];
                cplx b = transB == kNoTrans ? B[(size_t)k * ldb + j] : B[(size_t)j * ldb + k];
                if (transA == kConjTrans) a = std::conj(a);
                if (transB == kConjTrans) b = std::conj(b);
                acc += a * b;
            }
            C[(size_t)i * ldc + j] = alpha * acc + (beta == cplx(0.0) ? cplx(0.0) : beta * C[(size_t)i * ldc + j]);
        }
}

// A (m x n, row-major, m == n in the reference's only call) = U diag(s) VT, s descending.
// One-sided Jacobi (Hestenes) on the columns of A.
int LAPACKE_zgesvd(int layout, char jobu, char jobvt, int m, int n, cplx* a, int lda, double* s, cplx* u,
                   int ldu, cplx* vt, int ldvt, cplx* work, int lwork, double* rwork, int* info) {
    (void)layout; (void)jobu; (void)jobvt; (void)work; (void)lwork; (void)rwork;
    std::vector<cplx> W((size_t)m * n), V((size_t)n * n, cplx(0.0));
    for (int i = 0; i < m; i++)
        for (int j = 0; j < n; j++) W[(size_t)i * n + j] = a[(size_t)i * lda + j];
    for (int j = 0; j < n; j++) V[(size_t)j * n + j] = 1.0;
    for (int sweep = 0; sweep < 80; sweep++) {
        double off = 0.0;
        for (int p = 0; p < n; p++)
            for (int q = p + 1; q < n; q++) {
                double app = 0.0, aqq = 0.0;
                cplx apq = 0.0;
                for (int i = 0; i < m; i++) {
                    app += std::norm(W[(size_t)i * n + p]);
                    aqq += std::norm(W[(size_t)i * n + q]);
                    apq += std::conj(W[(size_t)i * n + p]) * W[(size_t)i * n + q];
                }
                const double ab = std::abs(apq);
                if (ab <= 1e-17 * std::sqrt(app * aqq) || ab == 0.0) continue;
                off = std::max(off, ab / std::sqrt(app * aqq));
                const cplx ph = apq / ab;
                const double tau = (aqq - app) / (2.0 * ab);
                const double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
                const double c = 1.0 / std::sqrt(1.0 + t * t), sn = t * c;
                for (int i = 0; i < m; i++) {
                    const cplx wp = W[(size_t)i * n + p], wq = W[(size_t)i * n + q];
                    W[(size_t)i * n + p] = c * wp - sn * std::conj(ph) * wq;
                    W[(size_t)i * n + q] = sn * ph * wp + c * wq;
                }
                for (int i = 0; i < n; i++) {
                    const cplx vp = V[(size_t)i * n + p], vq = V[(size_t)i * n + q];
                    V[(size_t)i * n + p] = c * vp - sn * std::conj(ph) * vq;
                    V[(size_t)i * n + q] = sn * ph * vp + c * vq;
                }
            }
        if (off < 1e-15) break;
    }
    std::vector<double> sv(n);
    std::vector<int> order(n);
    for (int j = 0; j < n; j++) {
        double nn = 0.0;
        for (int i = 0; i < m; i++) nn += std::norm(W[(size_t)i * n + j]);
        sv[j] = std::sqrt(nn);
        order[j] = j;
    }
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++)
            if (sv[order[j]] > sv[order[i]]) std::swap(order[i], order[j]);
    for (int k = 0; k < n; k++) {
        const int j = order[k];
        s[k] = sv[j];
        for (int i = 0; i < m; i++) u[(size_t)i * ldu + k] = sv[j] > 0 ? W[(size_t)i * n + j] / sv[j] : cplx(i == k ? 1.0 : 0.0);
        for (int i = 0; i < n; i++) vt[(size_t)k * ldvt + i] = std::conj(V[(size_t)i * n + j]);
    }
    if (info) *info = 0;
    return 0;
}

}  // extern "C"
