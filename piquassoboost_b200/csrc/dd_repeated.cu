// dd_repeated.cu -- extended-precision evaluation of the repeated-row BBFG / Glynn sum
// (reference: BBFGPermanentCalculatorRepeated.hpp:116-307 with its n-ary reflected Gray counter,
// common/source/n_aryGrayCodeCounter.cpp:114-240): EVERY floating-point operation -- column-sum updates, the
// N-fold complex products, the binomial-weighted accumulation -- in double-double arithmetic (~106 bits).
//
// Role: the engine's counterpart of the reference's extended-precision repeated calculators
// (BBFGPermanentCalculatorRepeated::calculate(..., use_extended=true): long double; lib 9 "Inf": MPFR) and
// the companion of dd_plain.cu.  Straightforward by design (speed is not its purpose): every thread owns a
// contiguous range of the mixed-radix counter, seeds its Gray digits from the range start exactly as
// n_aryGrayCodeCounter::initialize does, and then walks the reflected code one step at a time.
#include <cuda_runtime.h>
#include <stdint.h>

#include "dd.cuh"
#include "glynn_plain.cuh"
#include "launch.h"

namespace pqb200 {

constexpr int kDDRepMaxN = kDDRepCols;    // distinct product-side columns (multiplicities are powers)
constexpr int kDDRepMaxD = kDDRepDigits;  // Gray-side rows besides the anchor

struct RepDDParams {
    const double* mat;    // device, R x N interleaved complex; row 0 = anchor, rows 1..D = digit rows
    const double* binom;  // device, [D][binom_stride][2]: C(r_d, k) as a double-double (hi, lo)
    int binom_stride;
    int D, N;             // N: distinct columns
    unsigned char r[kDDRepMaxD];   // multiplicity of digit row d+1 (radix r[d] + 1)
    unsigned char cm[kDDRepMaxN];  // multiplicity of column j: its column sum enters the product cm[j] times
    unsigned long long begin, end;  // this launch's range of the counter
    unsigned long long per_thread;
    double* partials;
    unsigned int* counter;
    double* out;
};

// acc += x * y with the product formed error-free (TwoProd by FMA)
__device__ __forceinline__ void dd_add_prod(dd& acc, double x, double y) {
    dd pr;
    pr.hi = x * y;
    pr.lo = fma(x, y, -pr.hi);
    dd_add_dd(acc, pr);
}

__global__ void __launch_bounds__(kPlainThreads, 1) glynn_repeated_dd_kernel(const RepDDParams p) {
    const int D = p.D, N = p.N;
    cdd tot = cdd_zero();
    const unsigned long long tid = (unsigned long long)blockIdx.x * kPlainThreads + threadIdx.x;
    unsigned long long lo = p.begin + tid * p.per_thread;
    unsigned long long hi = lo + p.per_thread;
    if (hi > p.end) hi = p.end;
    if (lo < hi) {
        unsigned char cnt[kDDRepMaxD], k[kDDRepMaxD];
        dd cr[kDDRepMaxN], ci[kDDRepMaxN];
        // ---- counter digits of `lo` (digit 0 least significant) and their reflected Gray values
        {
            unsigned long long v = lo;
            for (int d = 0; d < D; d++) {
                const unsigned int radix = (unsigned int)p.r[d] + 1u;
                cnt[d] = (unsigned char)(v % radix);
                v /= radix;
            }
            int par = 0;
            for (int d = D - 1; d >= 0; d--) {
                k[d] = par ? (unsigned char)(p.r[d] - cnt[d]) : cnt[d];
                par ^= k[d] & 1;
            }
        }
        for (int j = 0; j < N; j++) {
            cr[j].hi = __ldg(p.mat + 2 * j);
            cr[j].lo = 0.0;
            ci[j].hi = __ldg(p.mat + 2 * j + 1);
            ci[j].lo = 0.0;
        }
        for (int d = 0; d < D; d++) {
            const double coef = (double)((int)p.r[d] - 2 * (int)k[d]);
            const double* row = p.mat + 2 * (size_t)(d + 1) * N;
            for (int j = 0; j < N; j++) {
                dd_add_prod(cr[j], coef, __ldg(row + 2 * j));  // error-free product (the integer
                dd_add_prod(ci[j], coef, __ldg(row + 2 * j + 1));  // coefficient is not a power of two)
            }
        }
        cdd acc = cdd_zero();
        for (unsigned long long t = lo; t < hi; t++) {
            // weight (a product of binomials, each given as an exact double-double) and sign of this term
            dd w;
            w.hi = 1.0;
            w.lo = 0.0;
            int ksum = 0;
            for (int d = 0; d < D; d++) {
                dd b;
                b.hi = __ldg(p.binom + 2 * (d * p.binom_stride + k[d]));
                b.lo = __ldg(p.binom + 2 * (d * p.binom_stride + k[d]) + 1);
                w = dd_mul_dd(w, b);
                ksum += k[d];
            }
            if (ksum & 1) w = dd_neg(w);
            // prod_j colsum_j^cm_j by repeated multiplication (BBFGPermanentCalculatorRepeated.hpp:263-275)
            cdd q;
            q.re = cr[0];
            q.im = ci[0];
            for (int j = 0; j < N; j++) {
                cdd f;
                f.re = cr[j];
                f.im = ci[j];
                for (int e = (j == 0 ? 1 : 0); e < (int)p.cm[j]; e++) q = cdd_mul(q, f);
            }
            q.re = dd_mul_dd(q.re, w);
            q.im = dd_mul_dd(q.im, w);
            cdd_add(acc, q);
            if (t + 1 < hi) {
                // n_aryGrayCodeCounter::next: increment the counter, recompute the reflected digits; exactly
                // one of them changes, by +-1
                int d = 0;
                while (cnt[d] == p.r[d]) {
                    cnt[d] = 0;
                    d++;
                }
                cnt[d]++;
                int par = 0, changed = -1, delta = 0;
                for (int e = D - 1; e >= 0; e--) {
                    const unsigned char g = par ? (unsigned char)(p.r[e] - cnt[e]) : cnt[e];
                    if (g != k[e]) {
                        changed = e;
                        delta = (int)g - (int)k[e];
                        k[e] = g;
                    }
                    par ^= g & 1;
                }
                const double sg = -2.0 * (double)delta;  // k -> k+delta: column sums change by -2*delta*a
                const double* row = p.mat + 2 * (size_t)(changed + 1) * N;
                for (int j = 0; j < N; j++) {
                    dd_add_d(cr[j], sg * __ldg(row + 2 * j));
                    dd_add_d(ci[j], sg * __ldg(row + 2 * j + 1));
                }
            }
        }
        tot = acc;
    }
    grid_reduce_cdd<kPlainThreads>(tot, p.partials, p.counter, p.out);
}

cudaError_t launch_repeated_dd(const double* d_mat, const double* d_binom, int binom_stride, int D, int N,
                               const int* r, const int* cm, unsigned long long begin, unsigned long long end, int max_blocks,
                               double* partials, unsigned int* counter, double* out, cudaStream_t stream) {
    RepDDParams p;
    p.mat = d_mat;
    p.binom = d_binom;
    p.D = D;
    p.N = N;
    p.binom_stride = binom_stride;
    for (int d = 0; d < kDDRepMaxD; d++) p.r[d] = (unsigned char)(d < D ? r[d] : 0);
    for (int j = 0; j < kDDRepMaxN; j++) p.cm[j] = (unsigned char)(j < N ? cm[j] : 1);
    p.begin = begin;
    p.end = end;
    const unsigned long long total = end - begin;
    unsigned long long per = 64;  // at least 64 steps per thread (the seeding costs ~D row updates)
    unsigned long long threads = (total + per - 1) / per;
    const unsigned long long max_threads = (unsigned long long)max_blocks * kPlainThreads;
    if (threads > max_threads) {
        per = (total + max_threads - 1) / max_threads;
        threads = (total + per - 1) / per;
    }
    p.per_thread = per;
    const int grid = (int)((threads + kPlainThreads - 1) / kPlainThreads);
    p.partials = partials;
    p.counter = counter;
    p.out = out;
    glynn_repeated_dd_kernel<<<grid < 1 ? 1 : grid, kPlainThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace pqb200
