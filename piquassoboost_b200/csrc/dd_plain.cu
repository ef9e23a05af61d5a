// dd_plain.cu -- extended-precision evaluation of the plain Glynn / BBFG sum: EVERY operation
// (column-sum updates, the n-fold complex products, all accumulations) is carried out in
// double-double arithmetic (~106 mantissa bits), on the caller's matrix as given (no normalisation).
//
// Role: the engine's counterpart of the reference's extended-precision calculators
// (BBFGPermanentCalculator::calculate(mtx, use_extended=true), BBFGPermanentCalculator.cpp:109-116:
// x87 long double, 64 mantissa bits) for sizes where the CPU reference is out of reach: it gives
// the n=40 ground truth the fast kernel is checked against (tests/test_plain_gpu.py, bench.py), and
// it is itself pinned against the reference's long double results for n <= 36.
// About 25x slower than glynn_plain_kernel; same Gray-code chunking, one chunk per thread.
#include <cuda_runtime.h>
#include <stdint.h>

#include "dd.cuh"
#include "glynn_plain.cuh"
#include "launch.h"

namespace pqb200 {

constexpr int kDDMaxN = 45;

struct PlainDDParams {
    const double* a;  // device, rows x n interleaved complex (row-major, dense)
    int rows;         // 2 <= rows <= n (rows < n: the caller padded the columns, see pqb200_permanent_dd)
    int n;
    int L;
    unsigned long long chunk_begin, chunk_end;
    unsigned int iters;
    double* partials;
    unsigned int* counter;
    double* out;
};

__device__ __forceinline__ void dd_acc_scaled(dd& acc, double x) { dd_add_d(acc, x); }

__global__ void __launch_bounds__(kPlainThreads, 1) glynn_plain_dd_kernel(const PlainDDParams p) {
    const int n = p.n, rows = p.rows;
    const int L = p.L;
    const unsigned int nsteps = 1u << L;
    const unsigned long long nthreads = (unsigned long long)gridDim.x * kPlainThreads;
    cdd tot = cdd_zero();
    unsigned long long c = p.chunk_begin + (unsigned long long)blockIdx.x * kPlainThreads + threadIdx.x;
    dd cr[kDDMaxN], ci[kDDMaxN];
    for (unsigned int it = 0; it < p.iters; it++, c += nthreads) {
        const bool live = c < p.chunk_end;
        const unsigned long long g0 = (live ? c : p.chunk_begin) << L;
        const unsigned long long gray0 = g0 ^ (g0 >> 1);
        for (int j = 0; j < n; j++) {
            cr[j].hi = __ldg(p.a + 2 * j);
            cr[j].lo = 0.0;
            ci[j].hi = __ldg(p.a + 2 * j + 1);
            ci[j].lo = 0.0;
        }
        for (int i = 1; i < rows; i++) {
            const double d = ((gray0 >> (rows - 1 - i)) & 1ull) ? -1.0 : 1.0;
            const double* row = p.a + 2 * (size_t)i * n;
            for (int j = 0; j < n; j++) {
                dd_add_d(cr[j], d * __ldg(row + 2 * j));
                dd_add_d(ci[j], d * __ldg(row + 2 * j + 1));
            }
        }
        cdd acc = cdd_zero();
        bool neg = (__popcll(gray0) & 1) != 0;
        for (unsigned int s = 0; s < nsteps; s++) {
            cdd q;
            q.re = cr[0];
            q.im = ci[0];
            for (int j = 1; j < n; j++) {
                cdd f;
                f.re = cr[j];
                f.im = ci[j];
                q = cdd_mul(q, f);
            }
            if (neg) {
                q.re = dd_neg(q.re);
                q.im = dd_neg(q.im);
            }
            cdd_add(acc, q);
            neg = !neg;
            // toggle into term g0+s+1
            const unsigned int s1 = s + 1;
            if (s1 < nsteps) {
                const int b = __ffs(s1) - 1;
                const unsigned long long g = g0 + s1;
                const double sg = (((g >> b) ^ (g >> (b + 1))) & 1ull) ? -2.0 : 2.0;
                const double* row = p.a + 2 * (size_t)(rows - 1 - b) * n;
                for (int j = 0; j < n; j++) {
                    dd_add_d(cr[j], sg * __ldg(row + 2 * j));
                    dd_add_d(ci[j], sg * __ldg(row + 2 * j + 1));
                }
            }
        }
        if (!live) acc = cdd_zero();
        cdd_add(tot, acc);
    }
    grid_reduce_cdd<kPlainThreads>(tot, p.partials, p.counter, p.out);
}

cudaError_t launch_plain_dd(const double* d_a, int rows, int n, int L, unsigned long long cb, unsigned long long ce,
                            int grid, double* partials, unsigned int* counter, double* out,
                            cudaStream_t stream) {
    PlainDDParams p;
    p.a = d_a;
    p.rows = rows;
    p.n = n;
    p.L = L;
    p.chunk_begin = cb;
    p.chunk_end = ce;
    const unsigned long long per_iter = (unsigned long long)grid * kPlainThreads;
    p.iters = (unsigned int)((ce - cb + per_iter - 1) / per_iter);
    p.partials = partials;
    p.counter = counter;
    p.out = out;
    glynn_plain_dd_kernel<<<grid, kPlainThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace pqb200
