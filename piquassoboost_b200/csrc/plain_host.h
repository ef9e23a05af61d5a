// plain_host.h -- host-side preparation of a matrix for glynn_plain_kernel.
//
// The permanent is multilinear in rows and in columns, so   perm(D_r A D_c) = prod(D_r) prod(D_c) perm(A)
// for diagonal D_r, D_c.  With unit-modulus factors chosen so that
//   * the LAST ROW   (toggled by the Gray code at every second step) is real and
//   * the LAST COLUMN is real in every row,
// the kernel saves N of the 2N additions of every fast toggle and 2 of the 4 multiplies of one
// factor per term (glynn_plain.cuh).  Unit-modulus factors leave every magnitude untouched: no
// over/underflow, zeros stay zeros, and each entry is rounded once more (relative 2^-53; measured
// effect on the permanent ~1e-15).  The factors are computed in long double and
// the inverse of their product is handed to the kernel as a double-double complex number.
//
// The same sum is what the reference evaluates (BBFGPermanentCalculator.hpp:107-232) also for
// rows != cols, where it is not multilinear in the rows: row factors are used for square input only;
// otherwise the last column is the constant padding column e_0 (column sum == 1), which is real.
#pragma once
#include <math.h>
#include <stddef.h>

#include <complex>
#include <vector>

namespace pqb200 {

// kernel column count for a rows x cols problem (2 <= rows <= cols+1)
inline int plain_kernel_cols(int rows, int cols) {
    return rows == cols ? cols : (rows > cols ? rows : cols + 1);
}

// a: rows x cols interleaved complex (stride in complex elements).  Writes the N x N (row-major,
// interleaved re,im) pre-doubled, normalised matrix in the kernel's row order to a2 (N = plain_kernel_cols(rows, cols)) and the
// dd complex correction factor to phase[4] = {re.hi, re.lo, im.hi, im.lo}.
// row stride (complex elements) of the kernel's matrix: rows that are a multiple of 256 bytes apart
// (N = 16, 32) are measurably slower through the constant path (all toggled rows alias), so those get
// one padding element
// (measured: N=32 26.5 -> 23.2 ms per 2^31 steps with the padding; no effect for other N)
#ifdef __CUDACC__
#define PQB_HOST_DEVICE __host__ __device__
#else
#define PQB_HOST_DEVICE
#endif
PQB_HOST_DEVICE constexpr int plain_row_stride(int n) { return n % 16 == 0 ? n + 1 : n; }

// ld: row stride of a2 in complex elements (0 = N)
inline void plain_normalise(const double* a, int rows, int cols, size_t stride, int N, double* a2,
                            double phase[4], int ld = 0) {
    if (ld == 0) ld = N;
    typedef std::complex<long double> cld;
    std::vector<cld> b((size_t)N * N, cld(0.0L, 0.0L));
    for (int i = 0; i < rows; i++)
        for (int j = 0; j < cols; j++)
            b[(size_t)i * N + j] = cld(a[2 * (i * stride + j)], a[2 * (i * stride + j) + 1]);
    for (int j = cols; j < N; j++) b[j] = cld(1.0L, 0.0L);  // padding columns: e_0, column sum == 1
    cld corr(1.0L, 0.0L);
    const int r = rows - 1;
    // column factors: make the last row real and non-negative
    for (int j = 0; j < cols; j++) {
        const cld v = b[(size_t)r * N + j];
        const long double m = hypotl(v.real(), v.imag());
        if (m == 0.0L || v.imag() == 0.0L) continue;
        const cld f = std::conj(v) / m;
        for (int i = 0; i < rows; i++) b[(size_t)i * N + j] *= f;
        b[(size_t)r * N + j] = cld(m, 0.0L);
        corr *= std::conj(f) / std::norm(f);  // 1/f
    }
    // row factors (square input only): make the last column real
    if (rows == cols) {
        for (int i = 0; i < r; i++) {
            const cld v = b[(size_t)i * N + (N - 1)];
            const long double m = hypotl(v.real(), v.imag());
            if (m == 0.0L || v.imag() == 0.0L) continue;
            const cld f = std::conj(v) / m;
            for (int j = 0; j < N; j++) b[(size_t)i * N + j] *= f;
            b[(size_t)i * N + (N - 1)] = cld(m, 0.0L);
            corr *= std::conj(f) / std::norm(f);
        }
    }
    // kernel row order: index k < rows-1 holds the row toggled by Gray bit k, i.e. matrix row
    // rows-1-k (index 0 = the real row); index rows-1 holds matrix row 0, whose delta is fixed to +1
    for (size_t k = 0; k < (size_t)2 * N * ld; k++) a2[k] = 0.0;
    for (int i = 0; i < rows; i++) {
        const int dst = i == 0 ? rows - 1 : rows - 1 - i;
        for (int j = 0; j < N; j++) {
            a2[2 * ((size_t)dst * ld + j)] = (double)(2.0L * b[(size_t)i * N + j].real());
            a2[2 * ((size_t)dst * ld + j) + 1] = (double)(2.0L * b[(size_t)i * N + j].imag());
        }
    }
    // the kernel relies on exact zeros in the imaginary parts of the last column ...
    for (int i = 0; i < N; i++) a2[2 * ((size_t)i * ld + (N - 1)) + 1] = 0.0;
    // ... and reads the real row (index 0) COMPACTLY: its N real parts are the first N doubles of the row
    if (rows >= 2) {
        std::vector<double> re((size_t)N);
        for (int j = 0; j < N; j++) re[j] = a2[2 * (size_t)j];
        for (int j = 0; j < 2 * N; j++) a2[j] = j < N ? re[j] : 0.0;
    }
    const double rh = (double)corr.real(), ih = (double)corr.imag();
    phase[0] = rh;
    phase[1] = (double)(corr.real() - (long double)rh);
    phase[2] = ih;
    phase[3] = (double)(corr.imag() - (long double)ih);
}

}  // namespace pqb200
