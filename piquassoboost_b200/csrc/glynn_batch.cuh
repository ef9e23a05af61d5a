// glynn_batch.cuh -- many small repeated-row permanents of sub-matrices of ONE matrix U in a single
// launch: the sub-permanent batches of the Clifford-Clifford sampler's pmf layer
// (reference: piquassoboost/sampling/simulation_strategies/source/samplingHelperFunctions.cpp:137-172,
// k calls of BBFGPermanentCalculatorRepeated per photon step per shot), the list-of-states form of
// the Python wrapper (sampling/ChinHuhPermanentCalculator_Wrapper.hpp:233-326) and
// BatchednPermanentCalculator (sampling/source/BatchedPermanentCalculator.cpp:109-296).
//
// One CTA per job.  A job names its rows and (physically repeated) columns by index into U, so no
// sub-matrix is ever built on the host (adaptInterferometer, BatchedPermanentCalculator.cpp:412-451,
// becomes the gather in the kernel prologue).  The CTA then runs the same low/high digit split as
// glynn_repeated.cuh, but builds the low-digit step table itself in shared memory (the table only
// depends on the job's radices) and reads the toggled row from shared memory: rows are uniform per
// CTA, not per launch, so they cannot live in the kernel-parameter bank.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dd.cuh"
#include "glynn_plain.cuh"
#include "glynn_repeated.cuh"

namespace pqb200 {

constexpr int kBatchThreads = 128;
constexpr int kBatchMaxRows = 32;    // rows of a job including the anchor
constexpr int kBatchMaxLen = 512;    // low-digit steps per chunk (shared-memory table)
constexpr int kBinomDim = 48;

// C(r,k) for r,k < kBinomDim; filled once per device by the host (exact in double)
static __constant__ double c_binom[kBinomDim][kBinomDim];

struct BatchJob {
    int R;                 // rows including the anchor (>= 2)
    int N;                 // physically repeated columns (= photons of the sub-problem)
    int q;                 // low digits
    unsigned int len;      // prod of low radices (<= kBatchMaxLen)
    unsigned int nchunks;  // prod of high radices
    int transpose;         // 1: element(g, p) = U[p][g] (Gray side indexes columns of U)
    int out_slot;
    int ustride;           // row stride (complex elements) of this job's matrix
    unsigned long long uoff;  // element offset of this job's matrix inside the U buffer
    unsigned char gidx[kBatchMaxRows];    // U index of each Gray-side row (row 0 = anchor)
    unsigned char radix[kBatchMaxRows];   // radix[d] = multiplicity of row d+1, plus 1
    unsigned char pidx[kBatchMaxRows + 16];  // U index of each product-side (expanded) column
};

template <int NMAX>
__global__ void __launch_bounds__(kBatchThreads)
glynn_batch_kernel(const double2* __restrict__ Ubuf, const BatchJob* __restrict__ jobs,
                   double2* __restrict__ out) {
    __shared__ double2 sA[kBatchMaxRows][NMAX];  // the job's sub-matrix (not doubled)
    __shared__ double2 sBase[2][NMAX];
    __shared__ RepStep sTable[kBatchMaxLen];
    __shared__ unsigned int sLowW[kBatchMaxRows], sHighW[kBatchMaxRows];
    __shared__ int sLowPar[2];
    __shared__ BatchJob sJob;
    __shared__ double sRed[kBatchThreads / 32][4];

    const int tid = threadIdx.x;
    {
        const int* src = reinterpret_cast<const int*>(jobs + blockIdx.x);
        int* dst = reinterpret_cast<int*>(&sJob);
        for (int i = tid; i < (int)(sizeof(BatchJob) / sizeof(int)); i += kBatchThreads) dst[i] = src[i];
    }
    __syncthreads();
    const int R = sJob.R, N = sJob.N, q = sJob.q, D = R - 1;
    const unsigned int len = sJob.len, nchunks = sJob.nchunks;
    const double2* U = Ubuf + sJob.uoff;
    const int ustride = sJob.ustride;
    // ---- gather the sub-matrix (adaptInterferometer + column repetition); pad columns with the
    //      constant 1 (anchor row 1, other rows 0) so that the product over NMAX columns is unchanged
    for (int e = tid; e < R * NMAX; e += kBatchThreads) {
        const int i = e / NMAX, j = e % NMAX;
        double2 v = make_double2(i == 0 ? 1.0 : 0.0, 0.0);
        if (j < N) {
            const int g = sJob.gidx[i], pcol = sJob.pidx[j];
            v = sJob.transpose ? U[(size_t)pcol * ustride + g] : U[(size_t)g * ustride + pcol];
        }
        sA[i][j] = v;
    }
    if (tid == 0) {
        unsigned int w = 1;
        for (int d = 0; d < q; d++) { sLowW[d] = w; w *= sJob.radix[d]; }
        w = 1;
        for (int d = q; d < D; d++) { sHighW[d] = w; w *= sJob.radix[d]; }
    }
    __syncthreads();
    // ---- low-digit step table (same content as repeated_host.h:rep_plan, built per CTA)
    for (unsigned int t = tid; t < len; t += kBatchThreads) {
        // changed digit: lowest digit with a non-zero counter value
        int ch = 0;
        for (int d = 0; d < q; d++) {
            if ((t / sLowW[d]) % sJob.radix[d] != 0) { ch = d; break; }
        }
        unsigned int flags = 0;
        double w = 1.0;
        for (int P = 0; P < 2; P++) {
            int par = P;
            for (int d = q - 1; d >= 0; d--) {
                const int r = sJob.radix[d] - 1;
                const int cnt = (int)((t / sLowW[d]) % sJob.radix[d]);
                const int g = par ? r - cnt : cnt;
                if (d == ch && par == 0) flags |= 1u << P;  // unmirrored digit: k increases with the counter
                par ^= g & 1;
                if (P == 0) w *= c_binom[r][g];
            }
        }
        RepStep e;
        e.w = (t & 1u) ? -w : w;
        e.row = ch + 1;
        e.flags = flags;
        sTable[t] = e;
    }
    // ---- base vectors and initial parities for both high parities
    for (int e = tid; e < 2 * NMAX; e += kBatchThreads) {
        const int P = e / NMAX, j = e % NMAX;
        double re = sA[0][j].x, im = sA[0][j].y;
        int par = P, ks = 0;
        for (int d = q - 1; d >= 0; d--) {
            const int r = sJob.radix[d] - 1;
            const int g = par ? r : 0;
            par ^= g & 1;
            ks += g;
            re += (double)(r - 2 * g) * sA[d + 1][j].x;
            im += (double)(r - 2 * g) * sA[d + 1][j].y;
        }
        sBase[P][j] = make_double2(re, im);
        if (j == 0) sLowPar[P] = ks & 1;
    }
    __syncthreads();

    cdd tot = cdd_zero();
    const unsigned int iters = (nchunks + kBatchThreads - 1) / kBatchThreads;
    for (unsigned int it = 0; it < iters; it++) {
        const unsigned int c = it * kBatchThreads + tid;
        const bool live = c < nchunks;
        const unsigned int cc = live ? c : 0u;
        // pass 1 (integers only): parity of the high Gray digits, binomial weight, sum of k
        int par = 0, ksum = 0;
        double whigh = 1.0;
        for (int d = D - 1; d >= q; d--) {
            const int radix = sJob.radix[d];
            const int cnt = (int)((cc / sHighW[d]) % (unsigned int)radix);
            const int k = par ? radix - 1 - cnt : cnt;
            par ^= k & 1;
            ksum += k;
            whigh *= c_binom[radix - 1][k];
        }
        const int P = par;
        double cr[NMAX], ci[NMAX];
#pragma unroll
        for (int j = 0; j < NMAX; j++) {
            cr[j] = sBase[P][j].x;
            ci[j] = sBase[P][j].y;
        }
        // pass 2: add the high rows
        par = 0;
        for (int d = D - 1; d >= q; d--) {
            const int radix = sJob.radix[d];
            const int cnt = (int)((cc / sHighW[d]) % (unsigned int)radix);
            const int k = par ? radix - 1 - cnt : cnt;
            par ^= k & 1;
            const double coef = (double)(radix - 1 - 2 * k);
#pragma unroll
            for (int j = 0; j < NMAX; j++) {
                const double2 a = sA[d + 1][j];
                cr[j] = fma(coef, a.x, cr[j]);
                ci[j] = fma(coef, a.y, ci[j]);
            }
        }
        cdd cacc = cdd_zero();
        double bs_re, bs_im;
        {
            double pr, pi;
            term_prod<NMAX>(cr, ci, pr, pi);
            const double w = sTable[0].w;
            bs_re = w * pr;
            bs_im = w * pi;
        }
        for (unsigned int t = 1; t < len; t++) {
            const RepStep e = sTable[t];
            const double sg = ((e.flags >> P) & 1u) ? -2.0 : 2.0;
#pragma unroll
            for (int j = 0; j < NMAX; j++) {
                const double2 a = sA[e.row][j];
                cr[j] = fma(sg, a.x, cr[j]);
                ci[j] = fma(sg, a.y, ci[j]);
            }
            double pr, pi;
            term_prod<NMAX>(cr, ci, pr, pi);
            bs_re = fma(e.w, pr, bs_re);
            bs_im = fma(e.w, pi, bs_im);
            if ((t & 7u) == 7u) {
                dd_add_d(cacc.re, bs_re);
                dd_add_d(cacc.im, bs_im);
                bs_re = 0.0;
                bs_im = 0.0;
            }
        }
        dd_add_d(cacc.re, bs_re);
        dd_add_d(cacc.im, bs_im);
        const double f = (((ksum + sLowPar[P]) & 1) ? -whigh : whigh) * (live ? 1.0 : 0.0);
        cdd scaled;
        scaled.re = dd_mul_d(cacc.re, f);
        scaled.im = dd_mul_d(cacc.im, f);
        cdd_add(tot, scaled);
    }
    // ---- CTA reduction (fixed order) and the exact 2^-(N-1) scaling
    tot = warp_reduce_cdd(tot);
    const int lane = tid & 31, warp = tid >> 5;
    if (lane == 0) {
        sRed[warp][0] = tot.re.hi;
        sRed[warp][1] = tot.re.lo;
        sRed[warp][2] = tot.im.hi;
        sRed[warp][3] = tot.im.lo;
    }
    __syncthreads();
    if (tid == 0) {
        cdd b = cdd_zero();
        for (int w = 0; w < kBatchThreads / 32; w++) {
            cdd x;
            x.re.hi = sRed[w][0];
            x.re.lo = sRed[w][1];
            x.im.hi = sRed[w][2];
            x.im.lo = sRed[w][3];
            cdd_add(b, x);
        }
        const double sc = ldexp(1.0, -(N - 1));
        out[sJob.out_slot] = make_double2((b.re.hi + b.re.lo) * sc, (b.im.hi + b.im.lo) * sc);
    }
}

// pmf[s][k] = | sum_i w[s][i] * addend[s][i] * U[k][i] |^2 / sum_k(...)   (samplingHelperFunctions.cpp:308-339)
// one CTA per shot; addend/weights are m-long per shot with zero weight where there is no photon
__global__ void pmf_combine_kernel(const double2* __restrict__ U, int ustride, int m,
                                   const double2* __restrict__ addend, const double* __restrict__ weight,
                                   double* __restrict__ pmf) {
    extern __shared__ double s_p[];  // m values + scratch
    const int s = blockIdx.x;
    for (int k = threadIdx.x; k < m; k += blockDim.x) {
        double re = 0.0, im = 0.0;
        for (int i = 0; i < m; i++) {
            const double w = weight[(size_t)s * m + i];
            if (w == 0.0) continue;
            const double2 a = addend[(size_t)s * m + i];
            const double wr = w * a.x, wi = w * a.y;
            const double2 u = U[(size_t)k * ustride + i];
            re += wr * u.x - wi * u.y;
            im += wr * u.y + wi * u.x;
        }
        s_p[k] = re * re + im * im;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int k = 0; k < m; k++) tot += s_p[k];  // same summation order as the reference loop
        s_p[m] = tot;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < m; k += blockDim.x) pmf[(size_t)s * m + k] = s_p[k] / s_p[m];
}

}  // namespace pqb200
