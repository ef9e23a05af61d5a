// glynn_plain_inst.cu -- instantiates glynn_plain_kernel<N> for N in [PQB_NLO, PQB_NHI] and registers
// the launchers.  Compiled several times with different ranges (see Makefile) so that the ~45
// instantiations build in parallel.
#include "glynn_plain.cuh"
#include "launch.h"
#include "plain_host.h"

#ifndef PQB_NLO
#error "define PQB_NLO and PQB_NHI"
#endif

namespace pqb200 {

template <int N>
static auto plain_kernel_fn() {
    if constexpr (default_unroll(N) == 2)
        return &glynn_plain_rolled_kernel<N>;
    else
        return &glynn_plain_kernel<N>;
}

template <int N>
static cudaError_t launch_plain_n(const PlainLaunchArgs& x) {
    static_assert(sizeof(PlainParams<N>) <= 32764, "kernel parameter space exceeded");
    static_assert(plain_anchor_bytes<N>() + 8 * kPlainThreads * sizeof(double) + 1024 <= 227 * 1024,
                  "anchors do not fit the shared memory of one SM");
    // opt in to the large dynamic shared memory once per device (the attribute is per context)
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        e = cudaFuncSetAttribute(plain_kernel_fn<N>(), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)plain_anchor_bytes<N>());
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    PlainParams<N> p;
    const int rows = (int)x.rows, cols = (int)x.cols;
    static_assert(sizeof(p.A2) == sizeof(double) * 2 * N * plain_row_stride(N), "A2 layout");
    plain_normalise(x.a, rows, cols, x.stride, N, reinterpret_cast<double*>(&p.A2[0][0]), p.phase,
                    plain_row_stride(N));
    p.rows = rows;
    p.L = x.L;
    p.K = x.K;
    p.chunk_begin = x.chunk_begin;
    p.chunk_end = x.chunk_end;
    const unsigned long long per_iter = (unsigned long long)x.grid * kPlainThreads;
    p.iters = (unsigned int)((x.chunk_end - x.chunk_begin + per_iter - 1) / per_iter);
    p.partials = x.partials;
    p.counter = x.counter;
    p.out = x.out;
    p.fast_row = 0;
    plain_kernel_fn<N>()<<<x.grid, kPlainThreads, plain_anchor_bytes<N>(), x.stream>>>(p);
    return cudaGetLastError();
}

template <int N>
static int occupancy_plain_n(void) {
    int nb = 0;
    cudaError_t e = cudaFuncSetAttribute(plain_kernel_fn<N>(), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)plain_anchor_bytes<N>());
    if (e == cudaSuccess)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, plain_kernel_fn<N>(), kPlainThreads,
                                                          plain_anchor_bytes<N>());
    return e == cudaSuccess ? nb : -1;
}

template <int LO, int HI>
struct FillPlain {
    static void go(plain_launch_fn* launch, plain_occupancy_fn* occ) {
        launch[LO] = &launch_plain_n<LO>;
        occ[LO] = &occupancy_plain_n<LO>;
        FillPlain<LO + 1, HI>::go(launch, occ);
    }
};
template <int HI>
struct FillPlain<HI, HI> {
    static void go(plain_launch_fn* launch, plain_occupancy_fn* occ) {
        launch[HI] = &launch_plain_n<HI>;
        occ[HI] = &occupancy_plain_n<HI>;
    }
};

}  // namespace pqb200

#define PQB_CAT2(a, b) a##b
#define PQB_CAT(a, b) PQB_CAT2(a, b)

extern "C" void PQB_CAT(pqb200_plain_fill_, PQB_NLO)(pqb200::plain_launch_fn* launch,
                                                     pqb200::plain_occupancy_fn* occ) {
    pqb200::FillPlain<PQB_NLO, PQB_NHI>::go(launch, occ);
}
