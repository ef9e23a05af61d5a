// glynn_reppow_inst.cu -- instantiates glynn_reppow_kernel<C> for C in [PQB_NLO, PQB_NHI] and
// registers the launchers (same scheme as glynn_plain_inst.cu).
#include "glynn_reppow.cuh"
#include "launch.h"
#include "repeated_host.h"

#ifndef PQB_NLO
#error "define PQB_NLO and PQB_NHI"
#endif

namespace pqb200 {

template <int C, bool GENERAL>
static cudaError_t launch_pow_cg(const PowLaunchArgs& x) {
    static_assert(sizeof(PowParams<C>) <= 32764, "kernel parameter space exceeded");
    constexpr int CS = pow_row_stride(C);
    if (x.C != C || x.R < 2 || x.R - 1 > kPowMaxDigits) return cudaErrorInvalidValue;
    if ((int)x.len > pow_len_max(x.R, C) || x.len < 1) return cudaErrorInvalidValue;
    static int occ[64] = {0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (occ[dev] == 0) {
        int nb = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, glynn_reppow_kernel<C, GENERAL>, kPlainThreads, 0);
        if (e != cudaSuccess) return e;
        occ[dev] = nb > 0 ? nb : 1;
    }
    PowParams<C> p;
    memset(&p, 0, sizeof(p));
    for (int i = 0; i < x.R; i++)
        for (int j = 0; j < C; j++)
            p.buf.c[i * CS + j] = make_double2(2.0 * x.mat[2 * ((size_t)i * C + j)],
                                               2.0 * x.mat[2 * ((size_t)i * C + j) + 1]);
    const int mat_doubles = 2 * x.R * CS;
    p.woff = mat_doubles;
    p.rowoff = 8 * (p.woff + (int)x.len + 1);
    p.flagoff = p.rowoff + (int)x.len + 2;
    if (p.flagoff + (int)x.len + 2 > kPowBufBytes) return cudaErrorInvalidValue;
    const RepStepHost* tab = (const RepStepHost*)x.table;
    for (unsigned int t = 0; t < x.len; t++) {
        p.buf.d[p.woff + t] = tab[t].w;
        p.buf.b[p.rowoff + t] = (unsigned char)tab[t].row;
        p.buf.b[p.flagoff + t] = (unsigned char)tab[t].flags;
    }
    p.buf.b[p.rowoff] = 1;
    p.buf.d[p.woff + x.len] = 0.0;
    p.buf.b[p.rowoff + x.len] = 1;  // dummy toggle behind the last term of a chunk (row 1, never used)
    p.buf.b[p.rowoff + x.len + 1] = 1;
    p.buf.b[p.flagoff + x.len] = 0;
    p.buf.b[p.flagoff + x.len + 1] = 0;
    for (int j = 0; j < kPowMaxC; j++) p.cm[j] = j < C ? x.cmult[j] : 1;
    p.D = x.R - 1;
    p.q = x.q;
    p.len = x.len;
    p.chunk_begin = x.chunk_begin;
    p.chunk_end = x.chunk_end;
    const unsigned long long want = (x.chunk_end - x.chunk_begin + kPlainThreads - 1) / kPlainThreads;
    const int grid = (int)std::min<unsigned long long>(
        std::max<unsigned long long>(want, 1), (unsigned long long)std::min(4096, occ[dev] * x.sm_count));
    const unsigned long long per_iter = (unsigned long long)grid * kPlainThreads;
    p.iters = (unsigned int)((x.chunk_end - x.chunk_begin + per_iter - 1) / per_iter);
    for (int d = 0; d < kPowMaxDigits; d++) {
        p.radix[d] = d < p.D ? x.radix[d] : 1;
        p.weight[d] = d < p.D ? x.weight[d] : 1;
        if (p.weight[d] == 0) p.weight[d] = 1;
    }
    p.real_row = 1;
    p.base = (const double2*)x.base;
    p.binom = x.binom;
    p.binom_stride = x.binom_stride;
    p.low_parity[0] = x.low_parity[0];
    p.low_parity[1] = x.low_parity[1];
    for (int k = 0; k < 4; k++) p.phase[k] = x.phase[k];
    p.partials = x.partials;
    p.counter = x.counter;
    p.out = x.out;
    glynn_reppow_kernel<C, GENERAL><<<grid, kPlainThreads, 0, x.stream>>>(p);
    return cudaGetLastError();
}

template <int C>
static cudaError_t launch_pow_c(const PowLaunchArgs& x) {
    bool general = false;
    for (int j = 0; j < x.C; j++) general = general || x.cmult[j] != 1;
    return general ? launch_pow_cg<C, true>(x) : launch_pow_cg<C, false>(x);
}

template <int LO, int HI>
struct FillPow {
    static void go(pow_launch_fn* launch) {
        launch[LO] = &launch_pow_c<LO>;
        FillPow<LO + 1, HI>::go(launch);
    }
};
template <int HI>
struct FillPow<HI, HI> {
    static void go(pow_launch_fn* launch) { launch[HI] = &launch_pow_c<HI>; }
};

}  // namespace pqb200

#define PQB_CAT2(a, b) a##b
#define PQB_CAT(a, b) PQB_CAT2(a, b)

extern "C" void PQB_CAT(pqb200_pow_fill_, PQB_NLO)(pqb200::pow_launch_fn* launch) {
    pqb200::FillPow<PQB_NLO, PQB_NHI>::go(launch);
}
