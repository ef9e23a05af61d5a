// glynn_plain.cuh -- the plain (no repeated rows) Glynn / BBFG Gray-code permanent kernel for sm_100a.
//
// What it computes (reference: piquassoboost/sampling/source/BBFGPermanentCalculator.hpp:107-232):
//
//     perm(A) = 2^-(n-1) * sum_{g=0}^{2^(n-1)-1}  parity(g) * prod_j colsum_j(g)
//     colsum_j(g) = a_0j + sum_{i>=1} delta_i(g) a_ij,  delta_i = -1 iff bit (n-1-i) of g^(g>>1) is set
//
// How it is laid out on the GPU (this is not the reference's TBB chunking):
//   * the 2^(n-1) Gray indices are cut into aligned chunks of 2^L steps; a THREAD owns a chunk, seeds
//     its N complex column sums from the chunk's start code (bounds rounding drift of the running
//     sums to 2^L updates) and keeps them in registers for the whole chunk;
//   * because chunks are 2^L aligned, every thread of the grid toggles the SAME row at the same
//     local step, so row data is a warp-uniform operand: the lowest S Gray bits ("static rows") are
//     unrolled with compile-time row index and sign and read 2*a_ij straight from the kernel
//     parameter constant bank (no load instruction at all); every 2^S steps one "runtime" row is
//     toggled from shared memory with a per-thread sign;
//   * per step: 2N DADD (column update) + 4(N-1) DMUL/DFMA (complex product chain, the last complex
//     multiply is fused into the block accumulator) = 6N-4 FP64 instructions for 8N algorithmic flops;
//   * per 2^S-term block the partial sum is folded into a double-double accumulator (TwoSum), chunk
//     sums are combined per thread, then warp shuffle -> shared -> one partial per CTA -> last CTA
//     reduces all partials in a fixed order (deterministic) into 4 doubles (re.hi re.lo im.hi im.lo).
//   * HBM traffic: 16*n*N bytes in, 32 bytes out. The kernel is bound by the FP64 pipe.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../piquassoboost_b200/csrc/dd.cuh"

namespace pqb200 {

constexpr int kPlainThreads = 256;

template <int N, int S>
struct PlainParams {
    double2 low[(1 << S) - 1][N];  // low[t-1][j] = -/+ 2*a[rows-1-ctz(t)][j]: the signed update applied before static term t
    const double2* A;   // device, rows x N row-major, the matrix itself
    int rows;           // n (>= S+1)
    int L;              // log2(chunk length); S <= L <= rows-1
    unsigned long long chunk_begin, chunk_end;  // this launch's (this GPU's) chunk range
    double* partials;       // [gridDim.x][4]
    unsigned int* counter;  // zero on entry, reset to zero on exit
    double* out;            // 4 doubles: re.hi re.lo im.hi im.lo  (unscaled sum over the chunk range)
};

__host__ __device__ constexpr int cx_ctz(int t) { return (t & 1) ? 0 : 1 + cx_ctz(t >> 1); }

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
__device__ __forceinline__ void grid_reduce_cdd(cdd tot, double* partials, unsigned int* counter,
                                                double* out) {
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

// One Gray term: acc += (NEG ? -1 : +1) * prod_j c[j], with CH independent multiply chains for ILP.
template <int N, bool NEG>
__device__ __forceinline__ void term_acc(const double (&cr)[N], const double (&ci)[N],
                                         double& acc_re, double& acc_im) {
    if (N == 1) {
        acc_re += NEG ? -cr[0] : cr[0];
        acc_im += NEG ? -ci[0] : ci[0];
        return;
    }
    constexpr int CH = (N >= 16) ? 4 : (N >= 4 ? 2 : 1);
    double qr[CH], qi[CH];
#pragma unroll
    for (int k = 0; k < CH; k++) {
        qr[k] = cr[k];
        qi[k] = ci[k];
    }
    // leave one factor out of the chains so that the last multiply can be fused into the accumulate
#pragma unroll
    for (int j = CH; j < N - 1; j++) cmul(qr[j % CH], qi[j % CH], cr[j], ci[j]);
    if (CH == 4) {
        cmul(qr[0], qi[0], qr[1], qi[1]);
        cmul(qr[2], qi[2], qr[3], qi[3]);
        cmul(qr[0], qi[0], qr[2], qi[2]);
    } else if (CH == 2) {
        cmul(qr[0], qi[0], qr[1], qi[1]);
    }
    if (N - 1 >= CH) {
        cmul_acc<NEG>(acc_re, acc_im, qr[0], qi[0], cr[N - 1], ci[N - 1]);
    } else {
        // N == CH (N==2 with CH... cannot happen: CH<=N/2 for N>=4, CH=1 for N<4)
        acc_re += NEG ? -qr[0] : qr[0];
        acc_im += NEG ? -qi[0] : qi[0];
    }
}

// The 2^S statically unrolled Gray terms t = T .. 2^S-1 of one block. Term t>0 first toggles static
// bit b = ctz(t) (row rows-1-b); the new bit value is bit b of t^(t>>1); parity alternates per term.
template <int N, int S, int T>
__device__ __forceinline__ void static_block(const PlainParams<N, S>& p, const unsigned int zoff,
                                             double (&cr)[N], double (&ci)[N], double& bs_re,
                                             double& bs_im) {
    if constexpr (T < (1 << S)) {
        if constexpr (T > 0) {
#pragma unroll
            for (int j = 0; j < N; j++) {
                cr[j] += p.low[T - 1][j + zoff].x;
                ci[j] += p.low[T - 1][j + zoff].y;
            }
        }
        term_acc<N, (T & 1) != 0>(cr, ci, bs_re, bs_im);
        static_block<N, S, T + 1>(p, zoff, cr, ci, bs_re, bs_im);
    }
}

template <int N, int S>
__global__ void __launch_bounds__(kPlainThreads, 1)
glynn_plain_kernel(const __grid_constant__ PlainParams<N, S> p) {
    extern __shared__ double2 sA[];  // rows x N
    const int rows = p.rows;
    for (int i = threadIdx.x; i < rows * N; i += kPlainThreads) sA[i] = p.A[i];
    __syncthreads();

    const int L = p.L;
    const unsigned int nblk = 1u << (L - S);
    const unsigned long long nthreads = (unsigned long long)gridDim.x * kPlainThreads;
    cdd tot = cdd_zero();

    for (unsigned long long c = p.chunk_begin + (unsigned long long)blockIdx.x * kPlainThreads + threadIdx.x;
         c < p.chunk_end; c += nthreads) {
        const unsigned long long g0 = c << L;
        const unsigned long long gray0 = g0 ^ (g0 >> 1);
        // ---- seed the column sums from the chunk's start code (reference: BBFG...hpp:136-176)
        double cr[N], ci[N];
#pragma unroll
        for (int j = 0; j < N; j++) {
            double2 a = sA[j];
            cr[j] = a.x;
            ci[j] = a.y;
        }
        for (int i = 1; i < rows; i++) {
            const double d = ((gray0 >> (rows - 1 - i)) & 1ull) ? -1.0 : 1.0;
#pragma unroll
            for (int j = 0; j < N; j++) {
                double2 a = sA[i * N + j];
                cr[j] = fma(d, a.x, cr[j]);
                ci[j] = fma(d, a.y, ci[j]);
            }
        }
        cdd cacc = cdd_zero();
        unsigned int zoff = 0;
        for (unsigned int k = 0; k < nblk; k++) {
            double bs_re = 0.0, bs_im = 0.0;
            // ---- 2^S terms with compile-time rows and signs (template recursion: straight-line code)
            // zoff is always 0; the empty asm makes it opaque per iteration so that the constant-bank
            // operands are re-read in place (LDCU right before use) instead of being hoisted out of
            // the loop into ~6N registers
            asm volatile("" : "+r"(zoff));
            static_block<N, S, 0>(p, zoff, cr, ci, bs_re, bs_im);
            dd_add_d(cacc.re, bs_re);
            dd_add_d(cacc.im, bs_im);
            // ---- one runtime toggle: Gray bit b = S + ctz(k+1), warp-uniform row, per-thread sign
            const unsigned int kk = k + 1;
            if (kk < nblk) {
                const int b = S + (__ffs(kk) - 1);
                const int row = rows - 1 - b;
                const unsigned long long g = g0 + ((unsigned long long)kk << S);
                const double sg = (((g >> b) ^ (g >> (b + 1))) & 1ull) ? -2.0 : 2.0;
                const double2* r = sA + row * N;
#pragma unroll
                for (int j = 0; j < N; j++) {
                    double2 a = r[j];
                    cr[j] = fma(sg, a.x, cr[j]);
                    ci[j] = fma(sg, a.y, ci[j]);
                }
            }
        }
        // chunk parity: every 2^S block starts with the parity of the chunk's start code
        if (__popcll(gray0) & 1) {
            cacc.re = dd_neg(cacc.re);
            cacc.im = dd_neg(cacc.im);
        }
        cdd_add(tot, cacc);
    }
    grid_reduce_cdd<kPlainThreads>(tot, p.partials, p.counter, p.out);
}

}  // namespace pqb200
