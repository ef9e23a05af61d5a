// experimental variant U: everything warp-uniform; matrix (2a) in the kernel-parameter constant bank, read with LDCU
#pragma once
#include "variant_static.cuh"
namespace pqb200 {
template <int N, int MAXROWS>
struct UParams {
    double2 A2[MAXROWS][N];   // 2a
    int rows, L;
    unsigned long long chunk_begin, chunk_end;
    unsigned int iters;       // uniform trip count of the chunk loop
    double* partials; unsigned int* counter; double* out;
};

template <int N, int MAXROWS>
__global__ void __launch_bounds__(kPlainThreads, 1)
glynn_u_kernel(const __grid_constant__ UParams<N, MAXROWS> p) {
    const int rows = p.rows;
    const int L = p.L;
    const unsigned int nsteps = 1u << L;
    const unsigned long long nthreads = (unsigned long long)gridDim.x * kPlainThreads;
    cdd tot = cdd_zero();
    unsigned long long c = p.chunk_begin + (unsigned long long)blockIdx.x * kPlainThreads + threadIdx.x;
    for (unsigned int it = 0; it < p.iters; it++, c += nthreads) {
        const bool live = c < p.chunk_end;
        const unsigned long long g0 = (live ? c : p.chunk_begin) << L;
        const unsigned long long gray0 = g0 ^ (g0 >> 1);
        double cr[N], ci[N];
#pragma unroll
        for (int j = 0; j < N; j++) { cr[j] = 0.5 * p.A2[0][j].x; ci[j] = 0.5 * p.A2[0][j].y; }
#pragma unroll 1
        for (int i = 1; i < rows; i++) {
            const double d = ((gray0 >> (rows - 1 - i)) & 1ull) ? -0.5 : 0.5;
#pragma unroll
            for (int j = 0; j < N; j++) { cr[j] = fma(d, p.A2[i][j].x, cr[j]); ci[j] = fma(d, p.A2[i][j].y, ci[j]); }
        }
        cdd cacc = cdd_zero();
#pragma unroll 1
        for (unsigned int s = 0; s < nsteps; s += 2) {
            double bs_re = 0.0, bs_im = 0.0;
            term_acc<N, false>(cr, ci, bs_re, bs_im);
            {
                const double sg = ((s >> 1) & 1u) ? 1.0 : -1.0;
#pragma unroll
                for (int j = 0; j < N; j++) { cr[j] = fma(sg, p.A2[rows - 1][j].x, cr[j]); ci[j] = fma(sg, p.A2[rows - 1][j].y, ci[j]); }
            }
            term_acc<N, true>(cr, ci, bs_re, bs_im);
            dd_add_d(cacc.re, bs_re);
            dd_add_d(cacc.im, bs_im);
            const unsigned int s2 = s + 2;
            if (s2 < nsteps) {
                const int b = __ffs(s2) - 1;
                const int row = rows - 1 - b;
                const unsigned long long g = g0 + s2;
                const double sg = (((g >> b) ^ (g >> (b + 1))) & 1ull) ? -1.0 : 1.0;
#pragma unroll
                for (int j = 0; j < N; j++) { cr[j] = fma(sg, p.A2[row][j].x, cr[j]); ci[j] = fma(sg, p.A2[row][j].y, ci[j]); }
            }
        }
        if (!live) cacc = cdd_zero();
        if (__popcll(gray0) & 1) { cacc.re = dd_neg(cacc.re); cacc.im = dd_neg(cacc.im); }
        cdd_add(tot, cacc);
    }
    grid_reduce_cdd<kPlainThreads>(tot, p.partials, p.counter, p.out);
}
}
