// glynn_plain.cuh -- the plain (no repeated rows) Glynn / BBFG Gray-code permanent kernel for sm_100a.
//
// What it computes (reference: piquassoboost/sampling/source/BBFGPermanentCalculator.hpp:107-232):
//
//     perm(A) = 2^-(n-1) * sum_{g=0}^{2^(n-1)-1}  parity(g) * prod_j colsum_j(g)
//     colsum_j(g) = a_0j + sum_{i>=1} delta_i(g) a_ij,  delta_i = -1 iff bit (n-1-i) of g^(g>>1) is set
//
// How it is laid out on the GPU (this is not the reference's TBB chunking):
//   * the 2^(n-1) Gray indices are cut into aligned chunks of 2^L steps; a THREAD owns a chunk, seeds
//     its N complex column sums from the chunk's start code (this bounds the rounding drift of the
//     running sums to 2^L updates) and keeps them in registers for the whole chunk;
//   * because chunks are 2^L aligned, every thread of the grid toggles the SAME row at the same
//     local step.  The whole matrix (pre-doubled, 2*a_ij) travels in the kernel-parameter constant
//     bank and the toggled row is read with warp-uniform constant loads (LDCU) straight into uniform
//     registers that feed the FP64 pipe as operands: no shared memory, no LSU traffic, no vector
//     registers spent on row data.  All loops have warp-uniform trip counts so that the compiler
//     can keep the row index in the uniform datapath; threads past the end of the chunk range run
//     a dummy chunk whose sum is discarded;
//   * per step: 2N DADD/DFMA (column update) + 4(N-1) DMUL/DFMA (complex product chains; the last
//     complex multiply is fused into the accumulator) = 6N-4 FP64 instructions for 8N algorithmic
//     flops;
//   * every DDP (default 8) terms the partial sum is folded into a double-double accumulator (TwoSum), chunk sums
//     are combined per thread in dd, then warp shuffle -> shared -> one partial per CTA -> the last
//     CTA reduces all partials in a fixed order (deterministic) into 4 doubles (re.hi re.lo im.hi
//     im.lo);
//   * HBM traffic: 16*n*N bytes of kernel parameters in, 32 bytes out.  Bound by the FP64 pipe.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dd.cuh"
#include "launch.h"

namespace pqb200 {

constexpr int kPlainMaxN = 45;  // 16*N*N + header must fit the 32764-byte kernel parameter space

template <int N>
struct PlainParams {
    double2 A2[N][N];  // 2*a[i][j] (exact doubling), rows x N used
    int rows;          // n, 2 <= rows <= N
    int L;             // log2(chunk length), 1 <= L <= rows-1
    unsigned long long chunk_begin, chunk_end;  // this launch's (this GPU's) chunk range
    unsigned int iters;                         // ceil((chunk_end-chunk_begin) / (gridDim.x*256))
    double* partials;                           // [gridDim.x][4]
    unsigned int* counter;                      // zero on entry, reset to zero on exit
    double* out;  // 4 doubles: re.hi re.lo im.hi im.lo  (unscaled sum over the chunk range)
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
// multiply chains per term, tuned on B200 (tools/microbench/mb2.cu, profiles/mb2_r1.log): long products
// want few chains (register pressure at N >= 36), short ones want many (latency)
__host__ __device__ constexpr int default_chains(int n) {
    return n >= 36 ? 2 : n >= 26 ? 4 : n >= 18 ? 8 : n >= 8 ? 4 : n >= 4 ? 2 : 1;
}

template <int N, bool NEG, int CH = default_chains(N)>
__device__ __forceinline__ void term_acc(const double (&cr)[N], const double (&ci)[N],
                                         double& acc_re, double& acc_im) {
    if (N == 1) {
        acc_re += NEG ? -cr[0] : cr[0];
        acc_im += NEG ? -ci[0] : ci[0];
        return;
    }
    double qr[CH], qi[CH];
#pragma unroll
    for (int k = 0; k < CH; k++) {
        qr[k] = cr[k];
        qi[k] = ci[k];
    }
    // one factor is left out of the chains so that the last multiply can be fused into the accumulate
#pragma unroll
    for (int j = CH; j < N - 1; j++) cmul(qr[j % CH], qi[j % CH], cr[j], ci[j]);
    // combine the chains pairwise
#pragma unroll
    for (int w = 1; w < CH; w *= 2) {
#pragma unroll
        for (int k = 0; k + w < CH; k += 2 * w) cmul(qr[k], qi[k], qr[k + w], qi[k + w]);
    }
    cmul_acc<NEG>(acc_re, acc_im, qr[0], qi[0], cr[N - 1], ci[N - 1]);
}

// c += sg * row   (sg = +-1 (or +-0.5 when seeding) in a register; row in uniform registers)
template <int N>
__device__ __forceinline__ void toggle_row(const PlainParams<N>& p, int row, double sg,
                                           double (&cr)[N], double (&ci)[N]) {
#pragma unroll
    for (int j = 0; j < N; j++) {
        cr[j] = fma(sg, p.A2[row][j].x, cr[j]);
        ci[j] = fma(sg, p.A2[row][j].y, ci[j]);
    }
}

// DDP: the partial sum is folded into the double-double accumulator every DDP terms (a power of two
// >= 2); CH: independent multiply chains per term.
template <int N, int DDP = 8, int CH = default_chains(N)>
__global__ void __launch_bounds__(kPlainThreads, 1)
glynn_plain_kernel(const __grid_constant__ PlainParams<N> p) {
    const int rows = p.rows;
    const int L = p.L;
    const unsigned int nsteps = 1u << L;
    const unsigned long long nthreads = (unsigned long long)gridDim.x * kPlainThreads;
    cdd tot = cdd_zero();
    unsigned long long c =
        p.chunk_begin + (unsigned long long)blockIdx.x * kPlainThreads + threadIdx.x;
    for (unsigned int it = 0; it < p.iters; it++, c += nthreads) {
        const bool live = c < p.chunk_end;
        const unsigned long long g0 = (live ? c : p.chunk_begin) << L;
        const unsigned long long gray0 = g0 ^ (g0 >> 1);
        // ---- seed the column sums from the chunk's start code (reference: BBFG...hpp:136-176)
        double cr[N], ci[N];
#pragma unroll
        for (int j = 0; j < N; j++) {
            cr[j] = 0.5 * p.A2[0][j].x;
            ci[j] = 0.5 * p.A2[0][j].y;
        }
#pragma unroll 1
        for (int i = 1; i < rows; i++) {
            const double d = ((gray0 >> (rows - 1 - i)) & 1ull) ? -0.5 : 0.5;
            toggle_row<N>(p, i, d, cr, ci);
        }
        cdd cacc = cdd_zero();
        const unsigned int g0lo = (unsigned int)g0;
        double bs_re = 0.0, bs_im = 0.0;
#pragma unroll 1
        for (unsigned int s = 0; s < nsteps; s += 2) {
            // two consecutive terms g, g+1 (g = g0+s even): Gray bit 0 is set going into g+1 iff bit 1
            // of g is clear (warp-uniform whenever L >= 2)
            term_acc<N, false, CH>(cr, ci, bs_re, bs_im);
            toggle_row<N>(p, rows - 1, (((g0lo | s) >> 1) & 1u) ? 1.0 : -1.0, cr, ci);
            term_acc<N, true, CH>(cr, ci, bs_re, bs_im);
            if (DDP == 2 || (s & (DDP - 2)) == (DDP - 2)) {
                dd_add_d(cacc.re, bs_re);
                dd_add_d(cacc.im, bs_im);
                bs_re = 0.0;
                bs_im = 0.0;
            }
            // ---- general toggle into term g+2: Gray bit b = ctz(s+2) >= 1, warp-uniform row, per-thread
            // sign (it depends on the chunk index only for b == L-1).  Unconditional so that the loop
            // body is one basic block: after the last pair b == L and row rows-1-L >= 0 is toggled
            // into column sums that are never used again.
            const unsigned int s2 = s + 2;
            const int b = __ffs(s2) - 1;
            const unsigned long long g = g0 + s2;
            const double sg = (((g >> b) ^ (g >> (b + 1))) & 1ull) ? -1.0 : 1.0;
            toggle_row<N>(p, rows - 1 - b, sg, cr, ci);
        }
        if (DDP > 2) {  // leftover when the chunk is shorter than DDP terms (bs is zero otherwise)
            dd_add_d(cacc.re, bs_re);
            dd_add_d(cacc.im, bs_im);
        }
        // every group of terms starts with the parity of the chunk's start code
        if (__popcll(gray0) & 1) {
            cacc.re = dd_neg(cacc.re);
            cacc.im = dd_neg(cacc.im);
        }
        if (!live) cacc = cdd_zero();
        cdd_add(tot, cacc);
    }
    grid_reduce_cdd<kPlainThreads>(tot, p.partials, p.counter, p.out);
}

}  // namespace pqb200
