// glynn_repeated.cuh -- repeated-row (n-ary Gray code) Glynn / BBFG permanent kernel for sm_100a.
//
// What it computes (reference: piquassoboost/sampling/source/BBFGPermanentCalculatorRepeated.hpp:116-307,
// n-ary counter: common/source/n_aryGrayCodeCounter.cpp:114-240).  After the reference's anchor-row
// step (hpp:78-94) the rows 1..D carry multiplicities r_i and
//
//   perm = 2^-(N-1) sum_{0<=k_i<=r_i} (-1)^{sum k} prod_i C(r_i,k_i) prod_j (a_0j + sum_i (r_i-2k_i) a_ij)
//
// where j runs over the PHYSICALLY REPEATED columns (the host expands column multiplicities, so
// the product chain is the plain kernel's).  The k vectors are visited in the reference's reflected
// mixed-radix Gray order: digit d (row d+1, radix r_d+1), digit 0 least significant.
//
// GPU layout.  The digit space is split into LOW digits 0..q-1 (swept inside a chunk, `len` =
// prod of their radices steps) and HIGH digits q..D-1 (one combination per chunk; a thread owns a
// chunk).  Three facts make the chunk loop as cheap as the plain kernel's:
//   1. which low digit changes at local step t depends on t only -> the toggled ROW is
//      warp-uniform and its data comes from the kernel-parameter constant bank (LDCU -> uniform
//      register operands), exactly as in glynn_plain.cuh;
//   2. the direction of the change depends on the thread only through one bit, the parity P of the
//      thread's high Gray digits: a digit is either mirrored (k -> r-k) for all steps or not at all.
//      The host tabulates both directions per step; the thread selects with P;
//   3. C(r,k) = C(r,r-k), so the binomial weight of the low digits is the same for mirrored and
//      unmirrored threads: it is tabulated per step, pre-multiplied with the alternating parity.
// The step table (signed weight, toggled row, directions) lives in the kernel-parameter constant bank
// next to the matrix, so that the row index is a uniform value and the row itself comes in through
// LDCU as a uniform-register operand (a table in global memory pulls the index -- and with it every
// matrix element -- through the vector path: measured 55 % FP64 pipe utilisation).  The chunk length
// is therefore bounded by what is left of the 32 KB parameter space: rep_len_max(N) steps.
// As in glynn_plain.cuh the host normalises the matrix with unit-modulus factors so that the row
// of the fastest digit (row 1) and the last column are real, and every term is fused with the toggle
// into the next one.  Per step: (N or 2N-1) DFMA update + 4(N-2) DMUL/DFMA product + 3 (weighted
// accumulate).  Accumulation is double-double as in the plain kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dd.cuh"
#include "glynn_plain.cuh"
#include "launch.h"

namespace pqb200 {

constexpr int kRepMaxN = 44;      // 16*N*N + header must fit the kernel parameter space
constexpr int kRepMaxDigits = 44;

// steps per chunk that fit the parameter space next to the N x N matrix (10 bytes per step)
__host__ __device__ constexpr int rep_len_max(int n) {
    return (32764 - 16 * n * plain_row_stride(n) - 704) / 10 > 4096
               ? 4096
               : (32764 - 16 * n * plain_row_stride(n) - 704) / 10;
}

// step-table entry and plain product used by the shared-memory kernels (glynn_batch.cuh,
// laplace_sampler.cuh), which build their tables per CTA
struct __align__(16) RepStep {
    double w;            // (-1)^t * prod_{low digits} C(r_d, k_d(t))
    int row;             // matrix row toggled going into step t (1..D)
    unsigned int flags;  // bit P: 1 if k increases (column sums -= 2a) for a thread with high parity P
};

__host__ __device__ constexpr int batch_chains(int n) {
    return n >= 36 ? 2 : n >= 26 ? 4 : n >= 18 ? 8 : n >= 8 ? 4 : n >= 4 ? 2 : 1;
}

// prod_j c[j]
template <int N>
__device__ __forceinline__ void term_prod(const double (&cr)[N], const double (&ci)[N], double& pr,
                                          double& pi) {
    constexpr int CH = batch_chains(N) <= N ? batch_chains(N) : 1;
    double qr[CH], qi[CH];
#pragma unroll
    for (int k = 0; k < CH; k++) {
        qr[k] = cr[k];
        qi[k] = ci[k];
    }
#pragma unroll
    for (int j = CH; j < N; j++) cmul(qr[j % CH], qi[j % CH], cr[j], ci[j]);
#pragma unroll
    for (int w = 1; w < CH; w *= 2) {
#pragma unroll
        for (int k = 0; k + w < CH; k += 2 * w) cmul(qr[k], qi[k], qr[k + w], qi[k + w]);
    }
    pr = qr[0];
    pi = qi[0];
}

template <int N>
struct RepParams {
    double2 A2[N][plain_row_stride(N)];  // 2*a, row 0 = anchor, rows 1..D = digit rows, columns physically repeated
    // per local step t of a chunk:  w[t] = (-1)^t * prod_{low digits} C(r_d, k_d(t));
    // trow[t] = row toggled going INTO step t (1..D); tflags[t] bit P: k increases (column sums -= 2a)
    // for a thread with high parity P.  Entry len is a harmless dummy (row 1).
    double w[rep_len_max(N) + 1];
    // row index and direction flags in SEPARATE arrays: the row index must stay a uniform value (it
    // addresses the row loads), the flags meet the per-thread parity P in vector code; one combined
    // entry drags the row index -- and with it every row load -- onto the vector path (measured)
    unsigned char trow[rep_len_max(N) + 2];
    unsigned char tflags[rep_len_max(N) + 2];
    int D;             // number of digits (rows - 1)
    int q;             // number of low digits
    unsigned int len;  // steps per chunk = prod_{d<q} radix[d]
    unsigned long long chunk_begin, chunk_end;
    unsigned int iters;
    unsigned char radix[kRepMaxDigits];      // r_d + 1
    unsigned long long weight[kRepMaxDigits];  // prod_{q<=l<d} radix[l]  (d >= q)
    const double2* base;    // [2][N]: a_0 + sum_{low d} (+-r_d) a_d for P = 0 / 1   (NOT doubled)
    const double* binom;    // [D][kRepMaxN+1]: C(r_d, k)
    int low_parity[2];      // parity of sum of the low digits' initial k for P = 0 / 1
    double* partials;
    unsigned int* counter;
    double* out;
    double phase[4];        // unit-modulus correction factor (dd complex) applied to the sum
    int real_row;           // == 1, as a run-time value: with a literal row index ptxas hoists the whole
                            // real row into uniform registers and pushes the step counter out of them
};

// One weighted term, acc += w * prod_j c[j], fused with the toggle of row `row` into the next term
// (see term_toggle in glynn_plain.cuh).  REALROW: the toggled row is real (row 1).
#ifndef PQB_REP_LAG
#define PQB_REP_LAG 4
#endif
// The toggle of column j is issued PQB_REP_LAG columns behind the chain multiply that last used it: the
// row index comes out of the step table (a uniform load), the row loads depend on it, and a warp issues in
// order -- the first toggle must not sit at the top of the term.
template <int N, int CH, bool REALROW>
__device__ __forceinline__ void rep_term_toggle(const RepParams<N>& p, int row, double sg, double w,
                                                double (&cr)[N], double (&ci)[N], double& acc_re,
                                                double& acc_im) {
    constexpr int M = N - 1;  // complex factors; column N-1 is real
    constexpr int C = CH < M ? CH : M;
    constexpr int LAG = PQB_REP_LAG;
    double qr[C], qi[C];
#pragma unroll
    for (int k = 0; k < C; k++) {
        qr[k] = cr[k];
        qi[k] = ci[k];
    }
#pragma unroll
    for (int j = C; j < M; j++) {
        cmul(qr[j % C], qi[j % C], cr[j], ci[j]);
        if (j - LAG >= 0) {
            cr[j - LAG] = fma(sg, p.A2[row][j - LAG].x, cr[j - LAG]);
            if (!REALROW) ci[j - LAG] = fma(sg, p.A2[row][j - LAG].y, ci[j - LAG]);
        }
    }
#pragma unroll
    for (int v = 1; v < C; v *= 2) {
#pragma unroll
        for (int k = 0; k + v < C; k += 2 * v) cmul(qr[k], qi[k], qr[k + v], qi[k + v]);
    }
    const double f = w * cr[N - 1];
#pragma unroll
    for (int j = (M - LAG > 0 ? M - LAG : 0); j < M; j++) {
        cr[j] = fma(sg, p.A2[row][j].x, cr[j]);
        if (!REALROW) ci[j] = fma(sg, p.A2[row][j].y, ci[j]);
    }
    cr[N - 1] = fma(sg, p.A2[row][N - 1].x, cr[N - 1]);
    acc_re = fma(qr[0], f, acc_re);
    acc_im = fma(qi[0], f, acc_im);
}

// measured on C4-shaped problems (tools/rep_cases.py): the chain count matters little (+-3 %)
__host__ __device__ constexpr int rep_chains(int n) { return n >= 36 ? 4 : n >= 4 ? 2 : 1; }

template <int N>
__device__ __forceinline__ void rep_toggle(const RepParams<N>& p, int row, double sg,
                                           double (&cr)[N], double (&ci)[N]) {
#pragma unroll
    for (int j = 0; j < N; j++) {
        cr[j] = fma(sg, p.A2[row][j].x, cr[j]);
        if (j < N - 1) ci[j] = fma(sg, p.A2[row][j].y, ci[j]);
    }
}

template <int N>
__global__ void __launch_bounds__(kPlainThreads, 1)
glynn_repeated_kernel(const __grid_constant__ RepParams<N> p) {
    constexpr int CH = rep_chains(N);
    const int D = p.D, q = p.q;
    const unsigned int len = p.len;
    const unsigned long long nthreads = (unsigned long long)gridDim.x * kPlainThreads;
    cdd tot = cdd_zero();
    unsigned long long c =
        p.chunk_begin + (unsigned long long)blockIdx.x * kPlainThreads + threadIdx.x;
    for (unsigned int it = 0; it < p.iters; it++, c += nthreads) {
        const bool live = c < p.chunk_end;
        const unsigned long long cc = live ? c : p.chunk_begin;
        // ---- decode the high digits (most significant first), build their reflected Gray values,
        //      the high part of the column sums, the binomial weight and the parity
        //      (n_aryGrayCodeCounter::initialize, n_aryGrayCodeCounter.cpp:114-140)
        // pass 1 (integers only): parity of the high Gray digits, their binomial weight, sum of k
        int par = 0, ksum = 0;
        double whigh = 1.0;
#pragma unroll 1
        for (int d = D - 1; d >= q; d--) {
            const int radix = p.radix[d];
            const int cnt = (int)((cc / p.weight[d]) % (unsigned long long)radix);
            const int k = par ? radix - 1 - cnt : cnt;
            par ^= k & 1;
            ksum += k;
            whigh *= p.binom[d * (kRepMaxN + 1) + k];
        }
        const int P = par;
        double cr[N], ci[N];
#pragma unroll
        for (int j = 0; j < N; j++) {
            const double2 b = p.base[P * N + j];
            cr[j] = b.x;
            if (j < N - 1) ci[j] = b.y;
        }
        // pass 2: add the high rows, (r - 2k) * a = 0.5 (r - 2k) * (2a)
        par = 0;
#pragma unroll 1
        for (int d = D - 1; d >= q; d--) {
            const int radix = p.radix[d];
            const int cnt = (int)((cc / p.weight[d]) % (unsigned long long)radix);
            const int k = par ? radix - 1 - cnt : cnt;
            par ^= k & 1;
            rep_toggle<N>(p, d + 1, 0.5 * (double)(radix - 1 - 2 * k), cr, ci);
        }
        // ---- sweep the low digits: term t, then the toggle into term t+1 (table index t+1)
        cdd cacc = cdd_zero();
        double bs_re = 0.0, bs_im = 0.0;
#pragma unroll 1
        for (unsigned int t = 0; t < len; t++) {
            const double w = p.w[t];
            const int row = (int)p.trow[t + 1];
            const double sg = ((p.tflags[t + 1] >> P) & 1u) ? -1.0 : 1.0;
            if (row == 1)
                rep_term_toggle<N, CH, true>(p, p.real_row, sg, w, cr, ci, bs_re, bs_im);
            else
                rep_term_toggle<N, CH, false>(p, row, sg, w, cr, ci, bs_re, bs_im);
            if ((t & 7u) == 7u) {  // fold into the double-double accumulator every 8 terms
                dd_add_d(cacc.re, bs_re);
                dd_add_d(cacc.im, bs_im);
                bs_re = 0.0;
                bs_im = 0.0;
            }
        }
        dd_add_d(cacc.re, bs_re);
        dd_add_d(cacc.im, bs_im);
        // ---- thread-constant factors: high binomials and the parity of the chunk's first term
        const double f = (((ksum + p.low_parity[P]) & 1) ? -whigh : whigh) * (live ? 1.0 : 0.0);
        cdd scaled;  // f is an exactly representable integer (< 2^53 for N <= 53)
        scaled.re = dd_mul_d(cacc.re, f);
        scaled.im = dd_mul_d(cacc.im, f);
        cdd_add(tot, scaled);
    }
    grid_reduce_cdd<kPlainThreads>(tot, p.partials, p.counter, p.out, p.phase);
}

}  // namespace pqb200
