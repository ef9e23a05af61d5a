// glynn_repeated.cuh -- pieces of the repeated-row (n-ary Gray code) Glynn / BBFG evaluation shared by
// the shared-memory kernels (glynn_batch.cuh: one CTA per small permanent; laplace_sampler.cuh: one
// cluster per shot): the step-table entry of the low-digit sweep and the plain product of a column-sum
// vector.  The grid-wide repeated kernel itself is glynn_reppow.cuh (power form, matrix and step table
// in the kernel-parameter bank).
//
// Reference: piquassoboost/sampling/source/BBFGPermanentCalculatorRepeated.hpp:116-307; n-ary counter
// common/source/n_aryGrayCodeCounter.cpp:114-240.  The digit space is split into LOW digits (swept inside
// a chunk; which digit changes at local step t depends on t only, its direction on the thread only
// through the parity P of the thread's high Gray digits, and C(r,k) = C(r,r-k) makes the binomial weight
// thread-independent) and HIGH digits (one combination per chunk).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dd.cuh"
#include "glynn_plain.cuh"
#include "launch.h"

namespace pqb200 {

// step-table entry and plain product used by the shared-memory kernels (glynn_batch.cuh,
// laplace_sampler.cuh), which build their tables per CTA
struct __align__(16) RepStep {
    double w;            // (-1)^t * prod_{low digits} C(r_d, k_d(t))
    int row;             // matrix row toggled going into step t (1..D)
    unsigned int flags;  // bit P: 1 if k increases (column sums -= 2a) for a thread with high parity P
};

__host__ __device__ constexpr int batch_chains(int n) {
    return n >= 36 ? 2 : n >= 26 ? 4 : n >= 18 ? 8 : n >= 8 ? 4 : n >= 4 ? 2 : 1;
}

// prod_j c[j]
template <int N>
__device__ __forceinline__ void term_prod(const double (&cr)[N], const double (&ci)[N], double& pr,
                                          double& pi) {
    constexpr int CH = batch_chains(N) <= N ? batch_chains(N) : 1;
    double qr[CH], qi[CH];
#pragma unroll
    for (int k = 0; k < CH; k++) {
        qr[k] = cr[k];
        qi[k] = ci[k];
    }
#pragma unroll
    for (int j = CH; j < N; j++) cmul(qr[j % CH], qi[j % CH], cr[j], ci[j]);
#pragma unroll
    for (int w = 1; w < CH; w *= 2) {
#pragma unroll
        for (int k = 0; k + w < CH; k += 2 * w) cmul(qr[k], qi[k], qr[k + w], qi[k + w]);
    }
    pr = qr[0];
    pi = qi[0];
}

}  // namespace pqb200
