// experimental variant: rolled loop, 2 steps per iteration, all rows from shared (MODE 0) or param constant (MODE 1)
#pragma once
#include "variant_static.cuh"
namespace pqb200 {
template <int N, int MAXROWS>
struct RolledParams {
    double2 A2c[MAXROWS][N];   // 2a (only MODE 1)
    const double2* A;  // a
    int rows, L;
    unsigned long long chunk_begin, chunk_end;
    double* partials; unsigned int* counter; double* out;
};

template <int N, int MODE, int MAXROWS>
__global__ void __launch_bounds__(kPlainThreads, 1)
glynn_rolled_kernel(const __grid_constant__ RolledParams<N, MAXROWS> p) {
    extern __shared__ double2 sA[];
    const int rows = p.rows;
    for (int i = threadIdx.x; i < rows * N; i += kPlainThreads) sA[i] = p.A[i];
    __syncthreads();
    const int L = p.L;
    const unsigned int nsteps = 1u << L;
    const unsigned long long nthreads = (unsigned long long)gridDim.x * kPlainThreads;
    cdd tot = cdd_zero();
    for (unsigned long long c = p.chunk_begin + (unsigned long long)blockIdx.x * kPlainThreads + threadIdx.x;
         c < p.chunk_end; c += nthreads) {
        const unsigned long long g0 = c << L;
        const unsigned long long gray0 = g0 ^ (g0 >> 1);
        double cr[N], ci[N];
#pragma unroll
        for (int j = 0; j < N; j++) { double2 a = sA[j]; cr[j] = a.x; ci[j] = a.y; }
        for (int i = 1; i < rows; i++) {
            const double d = ((gray0 >> (rows - 1 - i)) & 1ull) ? -1.0 : 1.0;
#pragma unroll
            for (int j = 0; j < N; j++) { double2 a = sA[i * N + j]; cr[j] = fma(d, a.x, cr[j]); ci[j] = fma(d, a.y, ci[j]); }
        }
        cdd cacc = cdd_zero();
#pragma unroll 1
        for (unsigned int s = 0; s < nsteps; s += 2) {
            double bs_re = 0.0, bs_im = 0.0;
            term_acc<N, false>(cr, ci, bs_re, bs_im);
            {
                const double sg = ((s >> 1) & 1u) ? 2.0 : -2.0;
                if (MODE == 0) {
                    const double2* r = sA + (rows - 1) * N;
#pragma unroll
                    for (int j = 0; j < N; j++) { double2 a = r[j]; cr[j] = fma(sg, a.x, cr[j]); ci[j] = fma(sg, a.y, ci[j]); }
                } else {
                    const double sh = 0.5 * sg;
#pragma unroll
                    for (int j = 0; j < N; j++) { cr[j] = fma(sh, p.A2c[rows - 1][j].x, cr[j]); ci[j] = fma(sh, p.A2c[rows - 1][j].y, ci[j]); }
                }
            }
            term_acc<N, true>(cr, ci, bs_re, bs_im);
            dd_add_d(cacc.re, bs_re);
            dd_add_d(cacc.im, bs_im);
            const unsigned int s2 = s + 2;
            if (s2 < nsteps) {
                const int b = __ffs(s2) - 1;
                const int row = rows - 1 - b;
                const unsigned long long g = g0 + s2;
                const double sg = (((g >> b) ^ (g >> (b + 1))) & 1ull) ? -2.0 : 2.0;
                if (MODE == 0) {
                    const double2* r = sA + row * N;
#pragma unroll
                    for (int j = 0; j < N; j++) { double2 a = r[j]; cr[j] = fma(sg, a.x, cr[j]); ci[j] = fma(sg, a.y, ci[j]); }
                } else {
                    const double sh = 0.5 * sg;
#pragma unroll
                    for (int j = 0; j < N; j++) { cr[j] = fma(sh, p.A2c[row][j].x, cr[j]); ci[j] = fma(sh, p.A2c[row][j].y, ci[j]); }
                }
            }
        }
        if (__popcll(gray0) & 1) { cacc.re = dd_neg(cacc.re); cacc.im = dd_neg(cacc.im); }
        cdd_add(tot, cacc);
    }
    grid_reduce_cdd<kPlainThreads>(tot, p.partials, p.counter, p.out);
}
}
