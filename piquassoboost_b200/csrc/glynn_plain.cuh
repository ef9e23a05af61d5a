// glynn_plain.cuh -- the plain (no repeated rows) Glynn / BBFG Gray-code permanent kernel for sm_100a.
//
// What it computes (reference: piquassoboost/sampling/source/BBFGPermanentCalculator.hpp:107-232):
//
//     perm(A) = 2^-(n-1) * sum_{g=0}^{2^(n-1)-1}  parity(g) * prod_j colsum_j(g)
//     colsum_j(g) = a_0j + sum_{i>=1} delta_i(g) a_ij,  delta_i = -1 iff bit (n-1-i) of g^(g>>1) is set
//
// How it is laid out on the GPU (this is not the reference's TBB chunking):
//   * the 2^(n-1) Gray indices are cut into aligned chunks of 2^L steps; a THREAD owns a chunk and
//     keeps its N complex column sums in registers.  The sums are fully re-seeded from the chunk's
//     start code and, every 2^K steps, re-derived from a per-thread anchor in shared memory ("two-level
//     re-seeding" below): rounding drift of the running sums is the dominant error term;
//   * because chunks are aligned, every thread of the grid toggles the SAME row at the same local
//     step.  The whole matrix (pre-doubled, 2*a_ij; row k = the row toggled by Gray bit k, the
//     delta=+1 row last) travels in the kernel-parameter constant bank and the toggled row is read
//     with warp-uniform constant loads (LDCU) straight into uniform registers that feed the FP64
//     pipe as operands: no LSU traffic, no vector registers spent on row data.  Loop counters, row
//     indices and directions are functions of the local step only, so that they stay in the uniform
//     datapath; threads past the end of the chunk range run a dummy chunk whose sum is discarded;
//   * the host normalises the matrix with unit-modulus row/column factors (plain_host.h) so that the
//     row of Gray bit 0 (toggled at every second step) and the LAST COLUMN are real: the fast toggle
//     then updates real parts only (N instead of 2N adds) and the last factor of the product is a real
//     number fused into the accumulate (2 DFMA).  The correction factor is applied once, in
//     double-double, by the CTA that finishes the grid reduction;
//   * per step: (N + 2N-1)/2 DFMA (column update, averaged over fast/general toggles) + 4(N-2) + 2
//     DMUL/DFMA (complex product chains) = 5.5N-6.5 FP64 instructions for 8N algorithmic flops
//     (a straight complex evaluation needs 6N-4).  Every term is fused with the toggle into the next
//     one (term_toggle) so that independent updates sit between the dependent multiplies;
//   * two loop forms, chosen per N by measurement (default_unroll / default_chains):
//     glynn_plain_kernel (8 or 4 terms per iteration, compile-time rows and directions) and
//     glynn_plain_rolled_kernel (a pair of terms per iteration, run-time row index);
//   * every 8 terms the partial sum is folded into a double-double accumulator (TwoSum); chunk sums
//     are combined per thread in dd, then warp shuffle -> shared -> one partial per CTA -> the last
//     CTA reduces all partials in a fixed order (deterministic) into 4 doubles (re.hi re.lo im.hi
//     im.lo);
//   * HBM traffic: 16*n*N bytes of kernel parameters in, 32 bytes out.  Bound by the FP64 pipe.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dd.cuh"
#include "launch.h"
#include "plain_host.h"

namespace pqb200 {

constexpr int kPlainMaxN = 45;  // 16*N*N + header must fit the 32764-byte kernel parameter space

template <int N>
struct PlainParams {
    double2 A2[N][plain_row_stride(N)];  // 2*a[i][j] (exact doubling), rows x N used
    int rows;          // n, 2 <= rows <= N
    int L;             // log2(chunk length), 1 <= L <= rows-1
    int K;             // log2(sub-chunk length), 1 <= K <= L: the column sums are re-derived from a
                       // per-thread anchor (shared memory) every 2^K steps
    unsigned long long chunk_begin, chunk_end;  // this launch's (this GPU's) chunk range
    unsigned int iters;                         // ceil((chunk_end-chunk_begin) / (gridDim.x*256))
    double* partials;                           // [gridDim.x][4]
    unsigned int* counter;                      // zero on entry, reset to zero on exit
    double* out;  // 4 doubles: re.hi re.lo im.hi im.lo  (unscaled sum over the chunk range)
    double phase[4];  // unit-modulus correction factor (re.hi re.lo im.hi im.lo) applied to the sum
    int fast_row;     // == 0 (index of the real row); a run-time value for the rolled kernel, whose
                      // schedule is better when ptxas cannot turn the row loads into immediates
};

__device__ __forceinline__ double shfl_down_d(double v, int off) {
    return __shfl_down_sync(0xffffffffu, v, off);
}

__device__ __forceinline__ cdd warp_reduce_cdd(cdd v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        cdd o;
        o.re.hi = shfl_down_d(v.re.hi, off);
        o.re.lo = shfl_down_d(v.re.lo, off);
        o.im.hi = shfl_down_d(v.im.hi, off);
        o.im.lo = shfl_down_d(v.im.lo, off);
        cdd_add(v, o);
    }
    return v;
}

// Block + grid reduction of one complex dd per thread; result written by the last CTA to out[0..3].
// Deterministic: every level sums in a fixed order.
template <int THREADS>
__device__ __noinline__ void grid_reduce_cdd(cdd tot, double* partials, unsigned int* counter,
                                             double* out, const double* phase = nullptr) {
    __shared__ double s_red[THREADS / 32][4];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    tot = warp_reduce_cdd(tot);
    if (lane == 0) {
        s_red[warp][0] = tot.re.hi;
        s_red[warp][1] = tot.re.lo;
        s_red[warp][2] = tot.im.hi;
        s_red[warp][3] = tot.im.lo;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        cdd b = cdd_zero();
        for (int w = 0; w < THREADS / 32; w++) {
            cdd x;
            x.re.hi = s_red[w][0];
            x.re.lo = s_red[w][1];
            x.im.hi = s_red[w][2];
            x.im.lo = s_red[w][3];
            cdd_add(b, x);
        }
        double* p = partials + 4 * (size_t)blockIdx.x;
        p[0] = b.re.hi;
        p[1] = b.re.lo;
        p[2] = b.im.hi;
        p[3] = b.im.lo;
        __threadfence();
        unsigned int done = atomicAdd(counter, 1u);
        s_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && warp == 0) {
        __threadfence();
        cdd b = cdd_zero();
        for (unsigned int i = lane; i < gridDim.x; i += 32) {
            const volatile double* p = partials + 4 * (size_t)i;
            cdd x;
            x.re.hi = p[0];
            x.re.lo = p[1];
            x.im.hi = p[2];
            x.im.lo = p[3];
            cdd_add(b, x);
        }
        b = warp_reduce_cdd(b);
        if (lane == 0) {
            if (phase) {
                cdd f;
                f.re.hi = phase[0];
                f.re.lo = phase[1];
                f.im.hi = phase[2];
                f.im.lo = phase[3];
                b = cdd_mul(b, f);
            }
            out[0] = b.re.hi;
            out[1] = b.re.lo;
            out[2] = b.im.hi;
            out[3] = b.im.lo;
            *counter = 0u;
        }
    }
}

// acc += sgn * (x * y), complex, with 4 DFMA (sgn is a compile-time +-1 folded into operand negation)
template <bool NEG>
__device__ __forceinline__ void cmul_acc(double& acc_re, double& acc_im, double xr, double xi,
                                         double yr, double yi) {
    if (!NEG) {
        acc_re = fma(xr, yr, acc_re);
        acc_re = fma(-xi, yi, acc_re);
        acc_im = fma(xr, yi, acc_im);
        acc_im = fma(xi, yr, acc_im);
    } else {
        acc_re = fma(-xr, yr, acc_re);
        acc_re = fma(xi, yi, acc_re);
        acc_im = fma(-xr, yi, acc_im);
        acc_im = fma(-xi, yr, acc_im);
    }
}

__device__ __forceinline__ void cmul(double& xr, double& xi, double yr, double yi) {
    double r = xr * yr;
    double i = xr * yi;
    r = fma(-xi, yi, r);
    i = fma(xi, yr, i);
    xr = r;
    xi = i;
}

// Loop form and multiply-chain count per N: the argmin of the measured sweep
// tools/microbench/tune_plain.cu on B200 (profiles/tune_plain_final_r1.log).  The FP64 pipe only stays fed
// if ptxas interleaves the dependent multiply chains with the independent column updates, and how
// well it does that depends on N (register pressure) -- so the choice is tabulated, not modelled.
//   unroll 2   : rolled kernel (glynn_plain_rolled_kernel), best when the column sums alone take
//                >= 136 registers;   unroll 4 / 8 : terms per iteration of glynn_plain_kernel.
__host__ __device__ constexpr int default_unroll(int n) { return n >= 27 ? 2 : n == 16 ? 4 : 8; }
__host__ __device__ constexpr int default_chains(int n) {
    //                    n = 0  1  2  3  4  5  6  7  8 ...                      (profiles/tune_plain_final_r1.log)
    constexpr int t[46] = {1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 2, 8, 3, 2, 8, 2, 6, 6, 4, 2, 8, 2, 4, 8, 6, 6, 8, 7, 8, 7, 6, 8, 8, 7, 8, 7, 2, 6, 4, 7, 6, 6, 3, 4, 5};
    return t[n < 0 ? 0 : n > 45 ? 45 : n];
}

// column update  x <- x + sign * a   with  SIGNMODE 0: sign = +1 and 1: sign = -1 at compile time (DADD
// with the matrix element as an immediate constant-bank operand), 2: sign = sg at run time (DFMA)
template <int SIGNMODE>
__device__ __forceinline__ double upd(double x, double a, double sg) {
    return SIGNMODE == 0 ? x + a : SIGNMODE == 1 ? x - a : fma(sg, a, x);
}

// Real part of element j of row `row`.  The real row (index 0, Gray bit 0) is stored COMPACTLY: its N
// real parts are the first N doubles of the row, so that one LDCU.128 brings two of them (the row is
// read at every second step; uniform-load bandwidth is a co-limiter of this kernel).
template <int N, bool REALROW>
__device__ __forceinline__ double row_re(const PlainParams<N>& p, int row, int j) {
    if (REALROW) return reinterpret_cast<const double*>(&p.A2[row][0])[j];
    return p.A2[row][j].x;
}

// One Gray term, acc += (NEG ? -1 : +1) * prod_j c[j], fused with the toggle of row `row` into the
// next term: column j is updated right after its last use in the product, so that the independent
// update instructions sit between the dependent multiplies of the CH chains (they hide the FP64
// latency; 8 warps per SM cannot).  Columns 0..N-2 are complex, column N-1 is real (ci[N-1] is never
// touched) and is fused into the accumulate.  REALROW: the toggled row is real (Gray bit 0).
template <int N, bool NEG, int CH, bool REALROW, int SIGNMODE>
__device__ __forceinline__ void term_toggle(const PlainParams<N>& p, int row, double sg,
                                            double (&cr)[N], double (&ci)[N], double& acc_re,
                                            double& acc_im) {
    constexpr int M = N - 1;  // complex factors
    constexpr int C = CH < M ? CH : M;
    double qr[C], qi[C];
#pragma unroll
    for (int k = 0; k < C; k++) {
        qr[k] = cr[k];
        qi[k] = ci[k];
        cr[k] = upd<SIGNMODE>(cr[k], row_re<N, REALROW>(p, row, k), sg);
        if (!REALROW) ci[k] = upd<SIGNMODE>(ci[k], p.A2[row][k].y, sg);
    }
#pragma unroll
    for (int j = C; j < M; j++) {
        cmul(qr[j % C], qi[j % C], cr[j], ci[j]);
        cr[j] = upd<SIGNMODE>(cr[j], row_re<N, REALROW>(p, row, j), sg);
        if (!REALROW) ci[j] = upd<SIGNMODE>(ci[j], p.A2[row][j].y, sg);
    }
#pragma unroll
    for (int w = 1; w < C; w *= 2) {
#pragma unroll
        for (int k = 0; k + w < C; k += 2 * w) cmul(qr[k], qi[k], qr[k + w], qi[k + w]);
    }
    const double f = NEG ? -cr[N - 1] : cr[N - 1];
    cr[N - 1] = upd<SIGNMODE>(cr[N - 1], row_re<N, REALROW>(p, row, N - 1), sg);
    acc_re = fma(qr[0], f, acc_re);
    acc_im = fma(qi[0], f, acc_im);
}

// c += sg * row (seeding and the rare mid-chunk correction); row >= 1 (row 0 is stored compactly)
template <int N>
__device__ __forceinline__ void toggle_row(const PlainParams<N>& p, int row, double sg,
                                           double (&cr)[N], double (&ci)[N]) {
#pragma unroll
    for (int j = 0; j < N; j++) {
        cr[j] = fma(sg, p.A2[row][j].x, cr[j]);
        if (j < N - 1) ci[j] = fma(sg, p.A2[row][j].y, ci[j]);
    }
}

// chunk accumulator (re.hi re.lo im.hi im.lo at stride kPlainThreads in shared memory) += (re, im)
__device__ __forceinline__ void fold_chunk_acc(double* acc, double re, double im) {
    dd a;
    a.hi = acc[0 * kPlainThreads];
    a.lo = acc[1 * kPlainThreads];
    dd_add_d(a, re);
    acc[0 * kPlainThreads] = a.hi;
    acc[1 * kPlainThreads] = a.lo;
    a.hi = acc[2 * kPlainThreads];
    a.lo = acc[3 * kPlainThreads];
    dd_add_d(a, im);
    acc[2 * kPlainThreads] = a.hi;
    acc[3 * kPlainThreads] = a.lo;
}

__host__ __device__ constexpr int ctz_c(int x) { return (x & 1) ? 0 : 1 + ctz_c(x >> 1); }

// +-1.0 assembled from integer bits: -1.0 with the sign flipped when `bit` is set.  `vzero` is always
// zero but thread-dependent as far as the compiler can tell: it keeps the factor in a VECTOR register
// while the step counter that `bit` comes from never leaves the uniform datapath.
__device__ __forceinline__ double pm_one(unsigned int bit, unsigned int vzero) {
    return __hiloint2double((int)((0xbff00000u ^ (bit << 31)) ^ vzero), 0);
}

// Terms T..U-1 of one unrolled group of U = 2^LOGU consecutive Gray terms starting at local step s
// (s is a multiple of U).  Which row a term toggles, and for most of them the direction, depend on
// T alone: those rows are read as immediate constant-bank operands with the sign folded into the
// instruction.  Only the last toggle of the group (Gray bit >= LOGU) has a run-time row.
template <int N, int CH, int U, int T>
struct UnrolledTerms {
    static __device__ __forceinline__ void run(const PlainParams<N>& p, unsigned int s,
                                               unsigned int vzero, double (&cr)[N], double (&ci)[N],
                                               double& acc_re, double& acc_im) {
        constexpr bool NEG = (T & 1) != 0;
        if constexpr ((T & 1) == 0) {
            // into an odd term: Gray bit 0 (real row 0) is set iff bit 1 of the even index is clear
            term_toggle<N, NEG, CH, true, ((T >> 1) & 1) ? 0 : 1>(p, 0, 0.0, cr, ci, acc_re, acc_im);
        } else if constexpr (T + 1 < U) {
            // into the even term s+T+1: Gray bit b = ctz(T+1) changes, set iff bit b+1 of it is clear
            constexpr int b = ctz_c(T + 1);
            if constexpr ((2 << b) < U) {
                term_toggle<N, NEG, CH, false, (((T + 1) >> (b + 1)) & 1) ? 0 : 1>(p, b, 0.0, cr, ci, acc_re, acc_im);
            } else {  // bit b+1 == LOGU of s+T+1 is bit LOGU of s
                const double sg = pm_one((s / U) & 1u, vzero);
                term_toggle<N, NEG, CH, false, 2>(p, b, sg, cr, ci, acc_re, acc_im);
            }
        } else {
            const unsigned int s2 = s + U;
            const int b = __ffs(s2) - 1;
            const double sg = pm_one((s2 >> (b + 1)) & 1u, vzero);
            term_toggle<N, NEG, CH, false, 2>(p, b, sg, cr, ci, acc_re, acc_im);
        }
        if constexpr (T + 1 < U) UnrolledTerms<N, CH, U, T + 1>::run(p, s, vzero, cr, ci, acc_re, acc_im);
    }
};

// ---- two-level re-seeding --------------------------------------------------------------------------
// Rounding drift of the running column sums grows with the number of in-place updates since they were
// last derived from exact data (measured: it, not the FP64 products, dominates the error of the
// result; profiles/accuracy_*_r1.jsonl).  A chunk of 2^L steps is therefore swept as 2^(L-K)
// sub-chunks: the thread keeps an ANCHOR = column sums at the sub-chunk start in shared memory
// ((2N-1) doubles per thread, stride kPlainThreads: conflict-free), which only ever receives the
// single high-row toggle between consecutive sub-chunks (Gray code of the sub-chunk index), and the
// working sums are re-derived from it every 2^K steps:
//     c_start(sc) = anchor(sc)  -  (sc odd ? 2 a_{K-1} : 0)        (Gray bit K-1 follows bit 0 of sc)
// so that no column sum is ever more than 2^(L-K) + 2^K + 2 roundings away from a full seed, at the
// cost of (2N-1) LDS + STS and 2(2N-1) DFMA per 2^K steps (~1 %).
extern __shared__ double s_anchor[];

template <int N>
__device__ __forceinline__ void anchor_store(double* anc, const double (&cr)[N], const double (&ci)[N]) {
#pragma unroll
    for (int j = 0; j < N; j++) {
        anc[(2 * j) * kPlainThreads] = cr[j];
        if (j < N - 1) anc[(2 * j + 1) * kPlainThreads] = ci[j];
    }
}

// anchor += sg * A2[row];  c = anchor + fix * A2[fixrow]
template <int N>
__device__ __forceinline__ void anchor_step(const PlainParams<N>& p, double* anc, int row, double sg,
                                            int fixrow, double fix, double (&cr)[N], double (&ci)[N]) {
#pragma unroll
    for (int j = 0; j < N; j++) {
        double x = anc[(2 * j) * kPlainThreads];
        x = fma(sg, p.A2[row][j].x, x);
        anc[(2 * j) * kPlainThreads] = x;
        cr[j] = fma(fix, p.A2[fixrow][j].x, x);
        if (j < N - 1) {
            double y = anc[(2 * j + 1) * kPlainThreads];
            y = fma(sg, p.A2[row][j].y, y);
            anc[(2 * j + 1) * kPlainThreads] = y;
            ci[j] = fma(fix, p.A2[fixrow][j].y, y);
        }
    }
}

// Start of sub-chunk t of a chunk whose first sub-chunk index is sc0: full seed for t == 0
// (reference: BBFGPermanentCalculator.hpp:136-176), anchor step otherwise.
// Row order in A2: index k < rows-1 is the row of Gray bit k, index rows-1 the delta=+1 row.
template <int N>
__device__ __forceinline__ void subchunk_start(const PlainParams<N>& p, double* anc, unsigned int t,
                                               unsigned int nsub, unsigned long long sc, int K,
                                               double (&cr)[N], double (&ci)[N]) {
    if (t == 0) {
        const int rows = p.rows;
        const unsigned long long g0 = sc << K;
        const unsigned long long gray0 = g0 ^ (g0 >> 1);
#pragma unroll
        for (int j = 0; j < N; j++) {
            cr[j] = 0.5 * p.A2[rows - 1][j].x;
            if (j < N - 1) ci[j] = 0.5 * p.A2[rows - 1][j].y;
        }
        {
            const double d = (gray0 & 1ull) ? -0.5 : 0.5;  // Gray bit 0: the compact real row
#pragma unroll
            for (int j = 0; j < N; j++) cr[j] = fma(d, row_re<N, true>(p, 0, j), cr[j]);
        }
#pragma unroll 1
        for (int k = 1; k < rows - 1; k++) {
            const double d = ((gray0 >> k) & 1ull) ? -0.5 : 0.5;
            toggle_row<N>(p, k, d, cr, ci);
        }
        if (nsub > 1) anchor_store<N>(anc, cr, ci);
    } else {
        // Gray bit K+b of the step index (bit b of the sub-chunk index's Gray code), b = ctz(t), becomes
        // set iff bit b+1 of sc is clear
        const int b = __ffs(t) - 1;
        const double sg = ((sc >> (b + 1)) & 1ull) ? 1.0 : -1.0;
        anchor_step<N>(p, anc, K + b, sg, K - 1, (sc & 1ull) ? -1.0 : 0.0, cr, ci);
    }
}

// CH: independent multiply chains per term; U: terms per iteration of the unrolled chunk loop (the
// partial sum is folded into the double-double accumulator once per iteration).
template <int N, int CH = default_chains(N), int U = default_unroll(N)>
__global__ void __launch_bounds__(kPlainThreads, 1)
glynn_plain_kernel(const __grid_constant__ PlainParams<N> p) {
    const int L = p.L, K = p.K;
    const unsigned int nsteps = 1u << K;        // steps per sub-chunk
    const unsigned int nsub = 1u << (L - K);    // sub-chunks per chunk
    const unsigned long long nthreads = (unsigned long long)gridDim.x * kPlainThreads;
    double* const anc = s_anchor + threadIdx.x;
    // The double-double accumulators are touched once per U terms (chunk sum) / once per chunk
    // (thread total): they live in shared memory, not in the register file the column sums need.
    __shared__ double s_acc[8][kPlainThreads];
    double* const acc = &s_acc[0][threadIdx.x];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k * kPlainThreads] = 0.0;
    unsigned long long c =
        p.chunk_begin + (unsigned long long)blockIdx.x * kPlainThreads + threadIdx.x;
    double cr[N], ci[N];
    for (unsigned int it = 0; it < p.iters; it++, c += nthreads) {
      const bool live = c < p.chunk_end;
      const unsigned long long sc0 = (live ? c : p.chunk_begin) << (L - K);
#pragma unroll 1
      for (unsigned int t = 0; t < nsub; t++) {
        const unsigned long long cc = sc0 + t;  // sub-chunk index
        subchunk_start<N>(p, anc, t, nsub, cc, K, cr, ci);
        const unsigned long long gray0 = (cc << K) ^ ((cc << K) >> 1);
        double bs_re = 0.0, bs_im = 0.0;
        // Every quantity that selects a row or a sign inside the loops is a function of the local step
        // only (warp-uniform).  The one thread-dependent direction -- the toggle of Gray bit K-1 in
        // the middle of the sub-chunk, which follows bit 0 of the sub-chunk index -- is first applied with the
        // even-chunk direction and then corrected for odd chunks (factor 2 or 0, no divergence); the
        // chunk is swept as two segments so that the correction sits between two loops.
        const double midfix = (cc & 1ull) ? 2.0 : 0.0;
        const unsigned int vzero = (unsigned int)(cc >> 63) << 31;  // == 0, see pm_one
        const unsigned int seg = nsteps >= 4u ? (nsteps >> 1) : nsteps;
        if (seg >= (unsigned int)U) {
#pragma unroll 1
            for (unsigned int s0 = 0; s0 < nsteps; s0 += seg) {
#pragma unroll 1
                for (unsigned int s = s0; s < s0 + seg; s += U) {
                    UnrolledTerms<N, CH, U, 0>::run(p, s, vzero, cr, ci, bs_re, bs_im);
                    fold_chunk_acc(acc, bs_re, bs_im);
                    bs_re = 0.0;
                    bs_im = 0.0;
                }
                if (s0 == 0) toggle_row<N>(p, K - 1, midfix, cr, ci);
            }
        } else {
            // short chunks (tiny matrices): one pair of terms per iteration, run-time rows and signs.
            // (After the last pair b == L <= rows-1 is toggled into column sums never used again.)
#pragma unroll 1
            for (unsigned int s0 = 0; s0 < nsteps; s0 += seg) {
#pragma unroll 1
                for (unsigned int s = s0; s < s0 + seg; s += 2) {
                    const double sg1 = pm_one((s >> 1) & 1u, vzero);
                    const unsigned int s2 = s + 2;
                    const int b = __ffs(s2) - 1;
                    const double sg = pm_one((s2 >> (b + 1)) & 1u, vzero);
                    term_toggle<N, false, 1, true, 2>(p, 0, sg1, cr, ci, bs_re, bs_im);
                    term_toggle<N, true, 1, false, 2>(p, b, sg, cr, ci, bs_re, bs_im);
                    fold_chunk_acc(acc, bs_re, bs_im);
                    bs_re = 0.0;
                    bs_im = 0.0;
                }
                if (s0 == 0 && seg < nsteps) toggle_row<N>(p, K - 1, midfix, cr, ci);
            }
        }
        // chunk sum -> thread total; every group of terms starts with the parity of the chunk's
        // start code
        cdd cacc, tot;
        cacc.re.hi = acc[0 * kPlainThreads];
        cacc.re.lo = acc[1 * kPlainThreads];
        cacc.im.hi = acc[2 * kPlainThreads];
        cacc.im.lo = acc[3 * kPlainThreads];
        tot.re.hi = acc[4 * kPlainThreads];
        tot.re.lo = acc[5 * kPlainThreads];
        tot.im.hi = acc[6 * kPlainThreads];
        tot.im.lo = acc[7 * kPlainThreads];
        if (__popcll(gray0) & 1) {
            cacc.re = dd_neg(cacc.re);
            cacc.im = dd_neg(cacc.im);
        }
        if (!live) cacc = cdd_zero();
        cdd_add(tot, cacc);
        acc[0 * kPlainThreads] = 0.0;
        acc[1 * kPlainThreads] = 0.0;
        acc[2 * kPlainThreads] = 0.0;
        acc[3 * kPlainThreads] = 0.0;
        acc[4 * kPlainThreads] = tot.re.hi;
        acc[5 * kPlainThreads] = tot.re.lo;
        acc[6 * kPlainThreads] = tot.im.hi;
        acc[7 * kPlainThreads] = tot.im.lo;
      }
    }
    cdd tot;
    tot.re.hi = acc[4 * kPlainThreads];
    tot.re.lo = acc[5 * kPlainThreads];
    tot.im.hi = acc[6 * kPlainThreads];
    tot.im.lo = acc[7 * kPlainThreads];
    grid_reduce_cdd<kPlainThreads>(tot, p.partials, p.counter, p.out, p.phase);
}

// Rolled form of the same sweep (one pair of terms per loop iteration, accumulators in registers,
// run-time row for the second toggle).  ptxas schedules this form better than the unrolled one when
// the column sums alone take ~160 registers (N >= 35); which form serves which N is decided by
// measurement (tools/microbench/tune_plain.cu -> default_unroll).
template <int N, int CH = default_chains(N), int MINB = 1>
__global__ void __launch_bounds__(kPlainThreads, MINB)
glynn_plain_rolled_kernel(const __grid_constant__ PlainParams<N> p) {
    const int L = p.L, K = p.K;
    const unsigned int nsteps = 1u << K;        // steps per sub-chunk
    const unsigned int nsub = 1u << (L - K);    // sub-chunks per chunk
    const unsigned long long nthreads = (unsigned long long)gridDim.x * kPlainThreads;
    double* const anc = s_anchor + threadIdx.x;
    cdd tot = cdd_zero();
    unsigned long long c =
        p.chunk_begin + (unsigned long long)blockIdx.x * kPlainThreads + threadIdx.x;
    double cr[N], ci[N];
    for (unsigned int it = 0; it < p.iters; it++, c += nthreads) {
      const bool live = c < p.chunk_end;
      const unsigned long long sc0 = (live ? c : p.chunk_begin) << (L - K);
#pragma unroll 1
      for (unsigned int t = 0; t < nsub; t++) {
        const unsigned long long cc = sc0 + t;  // sub-chunk index
        subchunk_start<N>(p, anc, t, nsub, cc, K, cr, ci);
        const unsigned long long gray0 = (cc << K) ^ ((cc << K) >> 1);
        cdd cacc = cdd_zero();
        double bs_re = 0.0, bs_im = 0.0;
        const double midfix = (cc & 1ull) ? 2.0 : 0.0;  // see glynn_plain_kernel
        const unsigned int half = nsteps >> 1;
        const int r0 = p.fast_row;
#pragma unroll 1
        for (unsigned int s = 0; s < nsteps; s += 2) {
            const double sg1 = ((s >> 1) & 1u) ? 1.0 : -1.0;
            const unsigned int s2 = s + 2;
            const int b = __ffs(s2) - 1;
            const double sg = ((s2 >> (b + 1)) & 1u) ? 1.0 : -1.0;
            term_toggle<N, false, CH, true, 2>(p, r0, sg1, cr, ci, bs_re, bs_im);
            term_toggle<N, true, CH, false, 2>(p, r0 + b, sg, cr, ci, bs_re, bs_im);
            if (s2 == half) toggle_row<N>(p, K - 1, midfix, cr, ci);
            if ((s & 6u) == 6u) {
                dd_add_d(cacc.re, bs_re);
                dd_add_d(cacc.im, bs_im);
                bs_re = 0.0;
                bs_im = 0.0;
            }
        }
        dd_add_d(cacc.re, bs_re);  // leftover of chunks shorter than 8 terms (zero otherwise)
        dd_add_d(cacc.im, bs_im);
        if (__popcll(gray0) & 1) {
            cacc.re = dd_neg(cacc.re);
            cacc.im = dd_neg(cacc.im);
        }
        if (!live) cacc = cdd_zero();
        cdd_add(tot, cacc);
      }
    }
    grid_reduce_cdd<kPlainThreads>(tot, p.partials, p.counter, p.out, p.phase);
}

// dynamic shared memory of both kernels: the per-thread anchors
template <int N>
constexpr size_t plain_anchor_bytes() {
    return sizeof(double) * (size_t)(2 * N - 1) * kPlainThreads;
}

}  // namespace pqb200
