// glynn_reppow.cuh -- repeated-row / repeated-column Glynn / BBFG permanent kernel for sm_100a in the
// reference's POWER FORM: column multiplicities are NOT expanded.
//
// What it computes (reference: piquassoboost/sampling/source/BBFGPermanentCalculatorRepeated.hpp:116-307,
// inner loop :263-275; n-ary counter: common/source/n_aryGrayCodeCounter.cpp:114-240).  After the
// reference's anchor-row step (hpp:78-94) the rows 1..D carry multiplicities r_i and the C distinct
// columns multiplicities c_j (sum c_j = N photons):
//
//   perm = 2^-(N-1) sum_{0<=k_i<=r_i} (-1)^{sum k} prod_i C(r_i,k_i) prod_j (a_0j + sum_i (r_i-2k_i) a_ij)^{c_j}
//
// Per Gray step: C column-sum updates (not N) and N-1 complex multiplies -- each column sum is raised
// to its multiplicity by repeated multiplication exactly as the reference does -- so the work no
// longer grows with the expanded width, the thread keeps 2C-1 doubles of column sums (more resident
// warps for small C) and the matrix in the kernel-parameter bank is R x C: there is no photon-number
// cap (60 photons in 10 modes is an 11 x 10 matrix).
//
// GPU layout (same decomposition as the plain kernel, glynn_plain.cuh):
//   * digits 0..q-1 ("low") are swept inside a chunk of len = prod radix steps, digits q..D-1 ("high")
//     select the chunk; a THREAD owns a chunk.  Which low digit changes at local step t depends on t
//     only, so the toggled row is warp-uniform and comes out of the kernel-parameter constant bank as
//     uniform-register operands (LDCU); the direction depends on the thread only through the parity P
//     of its high Gray digits and is tabulated for both; C(r,k) = C(r,r-k) makes the binomial weight of
//     the low digits thread-independent (tabulated per step, pre-multiplied with the alternating sign);
//   * matrix, step weights, toggled-row indices and direction flags share ONE buffer in the parameter
//     bank (the matrix takes 16*R*CS bytes, the tables the rest: pow_len_max(R, C) steps per chunk);
//   * the host normalises with unit-modulus factors so that row 1 (fastest digit) and the last column
//     are real, puts the column of the highest multiplicity last (its power is a chain of real
//     multiplies) and sorts the others by descending multiplicity: the power loops of a group of 4
//     columns then share one uniform trip count;
//   * besides the double-double sum every thread accumulates sum |term| (2 FP64 instructions per
//     step): the host turns it into an a-posteriori bound of the rounding error and re-runs
//     ill-conditioned inputs in the double-double mode (pqb200_capi.cu: repeated_locked) -- the
//     engine's counterpart of the reference's precision dispatch
//     (BBFGPermanentCalculatorRepeated.cpp:105-141).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dd.cuh"
#include "glynn_plain.cuh"
#include "launch.h"

namespace pqb200 {

template <int C>
struct PowParams {
    union {
        double d[kPowBufBytes / 8];
        double2 c[kPowBufBytes / 16];
        unsigned char b[kPowBufBytes];
    } buf;        // [0, R*CS) double2: 2*a (row 0 = anchor, rows 1..D = digit rows); then the tables
    int woff;     // index into buf.d of w[0];    w[t] = (-1)^t prod_{low digits} C(r_d, k_d(t)), t <= len
    int rowoff;   // index into buf.b of trow[0]; trow[t] = row toggled going INTO step t (1..D), t <= len+1
    int flagoff;  // index into buf.b of tflags[0]; bit P: k increases (sums -= 2a) for high parity P
    unsigned char cm[kPowMaxC];  // column multiplicities, >= 1; cm[0..C-2] descending, cm[C-1] the largest
    int D;             // number of digits (rows - 1)
    int q;             // number of low digits
    unsigned int len;  // steps per chunk = prod_{d<q} radix[d]
    unsigned long long chunk_begin, chunk_end;
    unsigned int iters;
    unsigned char radix[kPowMaxDigits];        // r_d + 1
    unsigned long long weight[kPowMaxDigits];  // prod_{q<=l<d} radix[l]  (d >= q)
    const double2* base;  // [2][C]: a_0 + sum_{low d} (+-r_d) a_d for P = 0 / 1   (NOT doubled)
    const double* binom;  // [D][binom_stride]: C(r_d, k)
    int binom_stride;
    int low_parity[2];    // parity of the sum of the low digits' initial k for P = 0 / 1
    double* partials;     // [gridDim.x][8]
    unsigned int* counter;
    double* out;          // re.hi re.lo im.hi im.lo, then sum |term| (scaled like the sum)
    double phase[4];      // unit-modulus correction factor (dd complex) applied to the sum
    int real_row;         // == 1 as a run-time value (keeps ptxas from hoisting the real row, see glynn_repeated.cuh)
};

__host__ __device__ constexpr int pow_group(int c) { return c - 1 >= 4 ? 4 : c - 1 >= 2 ? 2 : 1; }
// multiply chains of the straight-line form (measured on C4-shaped problems: the count matters little)
#ifndef PQB_POW_CHAINS
#define PQB_POW_CHAINS(c) ((c) >= 36 ? 4 : (c) >= 4 ? 2 : 1)
#endif
__host__ __device__ constexpr int pow_chains(int c) { return PQB_POW_CHAINS(c); }

template <int C>
__device__ __forceinline__ double2 pow_elem(const PowParams<C>& p, int row, int j) {
    return p.buf.c[row * pow_row_stride(C) + j];
}

#ifndef PQB_POW_LAG
#define PQB_POW_LAG 4
#endif

// One weighted term, acc += w * prod_j c[j], fused with the toggle of row `row` into the next term --
// the form for product sides WITHOUT repeated columns (every cm[j] == 1: the C4 shape, Gray side
// repeated, product side single): one straight-line block, CH multiply chains, the toggle of column
// j issued PQB_POW_LAG columns behind the chain multiply that last used it (the row index comes out
// of the step table, the row loads depend on it and a warp issues in order -- the first toggle must
// not sit at the top of the term).  Columns 0..C-2 are complex, column C-1 is real.
template <int C, bool REALROW>
__device__ __forceinline__ void pow_term_toggle_single(const PowParams<C>& p, int row, double sg, double w,
                                                       double (&cr)[C], double (&ci)[C], double& acc_re,
                                                       double& acc_im, double& acc_abs) {
    constexpr int M = C - 1;  // complex factors; column C-1 is real
    constexpr int CH0 = pow_chains(C);
    constexpr int CH = M == 0 ? 1 : (CH0 < M ? CH0 : M);
    constexpr int LAG = PQB_POW_LAG;
    double qr[CH], qi[CH];
    if constexpr (M == 0) {
        qr[0] = 1.0;
        qi[0] = 0.0;
    } else {
#pragma unroll
        for (int k = 0; k < CH; k++) {
            qr[k] = cr[k];
            qi[k] = ci[k];
        }
#pragma unroll
        for (int j = CH; j < M; j++) {
            cmul(qr[j % CH], qi[j % CH], cr[j], ci[j]);
            if (j - LAG >= 0) {
                const double2 a = pow_elem<C>(p, row, j - LAG);
                cr[j - LAG] = fma(sg, a.x, cr[j - LAG]);
                if (!REALROW) ci[j - LAG] = fma(sg, a.y, ci[j - LAG]);
            }
        }
#pragma unroll
        for (int v = 1; v < CH; v *= 2) {
#pragma unroll
            for (int k = 0; k + v < CH; k += 2 * v) cmul(qr[k], qi[k], qr[k + v], qi[k + v]);
        }
    }
    const double f = w * cr[C - 1];
    if constexpr (M > 0) {
#pragma unroll
        for (int j = (M - LAG > 0 ? M - LAG : 0); j < M; j++) {
            const double2 a = pow_elem<C>(p, row, j);  // the columns the chain loop has not toggled yet
            cr[j] = fma(sg, a.x, cr[j]);
            if (!REALROW) ci[j] = fma(sg, a.y, ci[j]);
        }
    }
    cr[C - 1] = fma(sg, pow_elem<C>(p, row, C - 1).x, cr[C - 1]);
    acc_re = fma(qr[0], f, acc_re);
    acc_im = fma(qi[0], f, acc_im);
#ifndef PQB_POW_NOABS
    acc_abs = fma(fabs(qr[0]) + fabs(qi[0]), fabs(f), acc_abs);
#endif
}

// The general form, acc += w * prod_j c[j]^cm[j]: the complex columns are handled in groups of G, one
// running product per position in the group (G independent multiply chains); column sum j is
// multiplied into its chain cm[j] times (BBFGPermanentCalculatorRepeated.hpp:263-275).  The columns
// are sorted by descending multiplicity, so the first column of a group carries the group's trip
// count and the others are predicated (uniformly) only where multiplicities differ.  The previous
// group's columns are toggled while the current group is multiplied (their old values are dead).
template <int C, bool REALROW>
__device__ __forceinline__ void pow_term_toggle_general(const PowParams<C>& p, int row, double sg, double w,
                                                        double (&cr)[C], double (&ci)[C], double& acc_re,
                                                        double& acc_im, double& acc_abs) {
    constexpr int M = C - 1;  // complex columns
    constexpr int G = pow_group(C);
    double qr[G], qi[G];
#pragma unroll
    for (int u = 0; u < G; u++) {
        qr[u] = u < M ? cr[u] : 1.0;
        qi[u] = u < M ? ci[u] : 0.0;
    }
#pragma unroll
    for (int g0 = 0; g0 < M; g0 += G) {
        const int mmax = p.cm[g0];  // the largest multiplicity of the group (columns are sorted)
        if (g0 > 0) {
#pragma unroll
            for (int u = 0; u < G; u++) {
                const int j = g0 - G + u;
                const double2 a = pow_elem<C>(p, row, j);
                cr[j] = fma(sg, a.x, cr[j]);
                if (!REALROW) ci[j] = fma(sg, a.y, ci[j]);
            }
        }
        // the columns of the group are multiplied in as often as the group's SMALLEST multiplicity without
        // any test (the usual case: equal multiplicities), the rest under uniform predicates (measured: -2 ..
        // -6 % against one predicated loop; unrolling the loop by 2 or 4 is slower)
        const int glast = g0 + G - 1 < M - 1 ? g0 + G - 1 : M - 1;
        const int mmin = p.cm[glast];
        int k = g0 == 0 ? 1 : 0;
#pragma unroll 1
        for (; k < mmin; k++) {
#pragma unroll
            for (int u = 0; u < G; u++)
                if (g0 + u < M) cmul(qr[u], qi[u], cr[g0 + u], ci[g0 + u]);
        }
#pragma unroll 1
        for (; k < mmax; k++) {
#pragma unroll
            for (int u = 0; u < G; u++)
                if (g0 + u < M) {
                    if (u == 0 || k < (int)p.cm[g0 + u]) cmul(qr[u], qi[u], cr[g0 + u], ci[g0 + u]);
                }
        }
    }
    // the last group's toggles
    if constexpr (M > 0) {
        constexpr int last0 = ((M - 1) / G) * G;
#pragma unroll
        for (int j = last0; j < M; j++) {
            const double2 a = pow_elem<C>(p, row, j);
            cr[j] = fma(sg, a.x, cr[j]);
            if (!REALROW) ci[j] = fma(sg, a.y, ci[j]);
        }
    }
    constexpr int GU = G < M ? G : (M > 0 ? M : 1);  // chains actually in use
#pragma unroll
    for (int v = 1; v < GU; v *= 2) {
#pragma unroll
        for (int k = 0; k + v < GU; k += 2 * v) cmul(qr[k], qi[k], qr[k + v], qi[k + v]);
    }
    // real last column: f = w * c^cm
    double f = w * cr[C - 1];
    const int ml = p.cm[C - 1];
#pragma unroll 1
    for (int k = 1; k < ml; k++) f *= cr[C - 1];
    cr[C - 1] = fma(sg, pow_elem<C>(p, row, C - 1).x, cr[C - 1]);
    acc_re = fma(qr[0], f, acc_re);
    acc_im = fma(qi[0], f, acc_im);
    acc_abs = fma(fabs(qr[0]) + fabs(qi[0]), fabs(f), acc_abs);
}

template <int C, bool GENERAL, bool REALROW>
__device__ __forceinline__ void pow_term_toggle(const PowParams<C>& p, int row, double sg, double w,
                                                double (&cr)[C], double (&ci)[C], double& acc_re,
                                                double& acc_im, double& acc_abs) {
    if constexpr (GENERAL)
        pow_term_toggle_general<C, REALROW>(p, row, sg, w, cr, ci, acc_re, acc_im, acc_abs);
    else
        pow_term_toggle_single<C, REALROW>(p, row, sg, w, cr, ci, acc_re, acc_im, acc_abs);
}

template <int C>
__device__ __forceinline__ void pow_toggle(const PowParams<C>& p, int row, double sg, double (&cr)[C],
                                           double (&ci)[C]) {
#pragma unroll
    for (int j = 0; j < C; j++) {
        const double2 a = pow_elem<C>(p, row, j);
        cr[j] = fma(sg, a.x, cr[j]);
        if (j < C - 1) ci[j] = fma(sg, a.y, ci[j]);
    }
}

// Block + grid reduction of one complex dd and one plain double per thread (fixed order at every
// level: deterministic); the last CTA writes out[0..4].
template <int THREADS>
__device__ __noinline__ void grid_reduce_cdd_abs(cdd tot, double ab, double* partials,
                                                 unsigned int* counter, double* out, const double* phase) {
    __shared__ double s_red[THREADS / 32][5];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    tot = warp_reduce_cdd(tot);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ab += shfl_down_d(ab, off);
    if (lane == 0) {
        s_red[warp][0] = tot.re.hi;
        s_red[warp][1] = tot.re.lo;
        s_red[warp][2] = tot.im.hi;
        s_red[warp][3] = tot.im.lo;
        s_red[warp][4] = ab;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        cdd b = cdd_zero();
        double a = 0.0;
        for (int w = 0; w < THREADS / 32; w++) {
            cdd x;
            x.re.hi = s_red[w][0];
            x.re.lo = s_red[w][1];
            x.im.hi = s_red[w][2];
            x.im.lo = s_red[w][3];
            cdd_add(b, x);
            a += s_red[w][4];
        }
        double* pp = partials + 8 * (size_t)blockIdx.x;
        pp[0] = b.re.hi;
        pp[1] = b.re.lo;
        pp[2] = b.im.hi;
        pp[3] = b.im.lo;
        pp[4] = a;
        __threadfence();
        unsigned int done = atomicAdd(counter, 1u);
        s_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && warp == 0) {
        __threadfence();
        cdd b = cdd_zero();
        double a = 0.0;
        for (unsigned int i = lane; i < gridDim.x; i += 32) {
            const volatile double* pp = partials + 8 * (size_t)i;
            cdd x;
            x.re.hi = pp[0];
            x.re.lo = pp[1];
            x.im.hi = pp[2];
            x.im.lo = pp[3];
            cdd_add(b, x);
            a += pp[4];
        }
        b = warp_reduce_cdd(b);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) a += shfl_down_d(a, off);
        if (lane == 0) {
            cdd f;
            f.re.hi = phase[0];
            f.re.lo = phase[1];
            f.im.hi = phase[2];
            f.im.lo = phase[3];
            b = cdd_mul(b, f);
            out[0] = b.re.hi;
            out[1] = b.re.lo;
            out[2] = b.im.hi;
            out[3] = b.im.lo;
            out[4] = a;
            *counter = 0u;
        }
    }
}

// GENERAL = false: every column multiplicity is 1 (straight-line term); true: the power form.
template <int C, bool GENERAL>
__global__ void __launch_bounds__(kPlainThreads, pow_min_blocks(C))
glynn_reppow_kernel(const __grid_constant__ PowParams<C> p) {
    const int D = p.D, q = p.q;
    const unsigned int len = p.len;
    const unsigned long long nthreads = (unsigned long long)gridDim.x * kPlainThreads;
    cdd tot = cdd_zero();
    double tot_abs = 0.0;
    unsigned long long c =
        p.chunk_begin + (unsigned long long)blockIdx.x * kPlainThreads + threadIdx.x;
    for (unsigned int it = 0; it < p.iters; it++, c += nthreads) {
        const bool live = c < p.chunk_end;
        const unsigned long long cc = live ? c : p.chunk_begin;
        // ---- decode the high digits (most significant first): reflected Gray values, their parity,
        //      binomial weight and sum (n_aryGrayCodeCounter::initialize, n_aryGrayCodeCounter.cpp:114-140)
        int par = 0, ksum = 0;
        double whigh = 1.0;
#pragma unroll 1
        for (int d = D - 1; d >= q; d--) {
            const int radix = p.radix[d];
            const int cnt = (int)((cc / p.weight[d]) % (unsigned long long)radix);
            const int k = par ? radix - 1 - cnt : cnt;
            par ^= k & 1;
            ksum += k;
            whigh *= p.binom[d * p.binom_stride + k];
        }
        const int P = par;
        double cr[C], ci[C];
#pragma unroll
        for (int j = 0; j < C; j++) {
            const double2 b = p.base[P * C + j];
            cr[j] = b.x;
            if (j < C - 1) ci[j] = b.y;
        }
        // second pass: add the high rows, (r - 2k) * a = 0.5 (r - 2k) * (2a)
        par = 0;
#pragma unroll 1
        for (int d = D - 1; d >= q; d--) {
            const int radix = p.radix[d];
            const int cnt = (int)((cc / p.weight[d]) % (unsigned long long)radix);
            const int k = par ? radix - 1 - cnt : cnt;
            par ^= k & 1;
            pow_toggle<C>(p, d + 1, 0.5 * (double)(radix - 1 - 2 * k), cr, ci);
        }
        // ---- sweep the low digits: term t, then the toggle into term t+1 (table index t+1)
        cdd cacc = cdd_zero();
        double bs_re = 0.0, bs_im = 0.0, bs_abs = 0.0;
#pragma unroll 1
        for (unsigned int t = 0; t < len; t++) {
            const double w = p.buf.d[p.woff + t];
            const int row = (int)p.buf.b[p.rowoff + t + 1];
            const double sg = ((p.buf.b[p.flagoff + t + 1] >> P) & 1u) ? -1.0 : 1.0;
            if (row == 1)
                pow_term_toggle<C, GENERAL, true>(p, p.real_row, sg, w, cr, ci, bs_re, bs_im, bs_abs);
            else
                pow_term_toggle<C, GENERAL, false>(p, row, sg, w, cr, ci, bs_re, bs_im, bs_abs);
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
        cdd scaled;
        scaled.re = dd_mul_d(cacc.re, f);
        scaled.im = dd_mul_d(cacc.im, f);
        cdd_add(tot, scaled);
        tot_abs = fma(bs_abs, fabs(f), tot_abs);
    }
    grid_reduce_cdd_abs<kPlainThreads>(tot, tot_abs, p.partials, p.counter, p.out, p.phase);
}

}  // namespace pqb200
