// glynn_repeated_inst.cu -- instantiates glynn_repeated_kernel<N> for N in [PQB_NLO, PQB_NHI] and
// registers the launchers (same scheme as glynn_plain_inst.cu).
#include "glynn_repeated.cuh"
#include "launch.h"
#include "repeated_host.h"

#ifndef PQB_NLO
#error "define PQB_NLO and PQB_NHI"
#endif

namespace pqb200 {

static_assert(kRepBinomStride == kRepMaxN + 1, "binomial table stride");

template <int N>
static cudaError_t launch_rep_n(const RepLaunchArgs& x) {
    static_assert(sizeof(RepParams<N>) <= 32764, "kernel parameter space exceeded");
    RepParams<N> p;
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) {
            double2 v = make_double2(0.0, 0.0);
            if (i < x.R && j < x.N) {
                v.x = 2.0 * x.mat[2 * ((size_t)i * x.N + j)];
                v.y = 2.0 * x.mat[2 * ((size_t)i * x.N + j) + 1];
            }
            p.A2[i][j] = v;
        }
    p.D = x.R - 1;
    p.q = x.q;
    p.len = x.len;
    p.chunk_begin = x.chunk_begin;
    p.chunk_end = x.chunk_end;
    const unsigned long long per_iter = (unsigned long long)x.grid * kPlainThreads;
    p.iters = (unsigned int)((x.chunk_end - x.chunk_begin + per_iter - 1) / per_iter);
    for (int d = 0; d < kRepMaxDigits; d++) {
        p.radix[d] = d < p.D ? x.radix[d] : 1;
        p.weight[d] = d < p.D ? x.weight[d] : 1;
        if (p.weight[d] == 0) p.weight[d] = 1;
    }
    if (x.len > (unsigned int)rep_len_max(N)) return cudaErrorInvalidValue;
    const RepStepHost* tab = (const RepStepHost*)x.table;
    for (unsigned int t = 0; t < x.len; t++) {
        p.w[t] = tab[t].w;
        p.trow[t] = (unsigned char)tab[t].row;
        p.tflags[t] = (unsigned char)tab[t].flags;
    }
    p.trow[0] = 1;
    p.w[x.len] = 0.0;
    p.trow[x.len] = 1;  // dummy toggle behind the last term of a chunk (row 1, never used)
    p.tflags[x.len] = 0;
    p.real_row = 1;
    p.base = (const double2*)x.base;
    p.binom = x.binom;
    p.low_parity[0] = x.low_parity[0];
    p.low_parity[1] = x.low_parity[1];
    for (int k = 0; k < 4; k++) p.phase[k] = x.phase[k];
    p.partials = x.partials;
    p.counter = x.counter;
    p.out = x.out;
    glynn_repeated_kernel<N><<<x.grid, kPlainThreads, 0, x.stream>>>(p);
    return cudaGetLastError();
}

template <int LO, int HI>
struct FillRep {
    static void go(rep_launch_fn* launch) {
        launch[LO] = &launch_rep_n<LO>;
        FillRep<LO + 1, HI>::go(launch);
    }
};
template <int HI>
struct FillRep<HI, HI> {
    static void go(rep_launch_fn* launch) { launch[HI] = &launch_rep_n<HI>; }
};

}  // namespace pqb200

#define PQB_CAT2(a, b) a##b
#define PQB_CAT(a, b) PQB_CAT2(a, b)

extern "C" void PQB_CAT(pqb200_rep_fill_, PQB_NLO)(pqb200::rep_launch_fn* launch) {
    pqb200::FillRep<PQB_NLO, PQB_NHI>::go(launch);
}
