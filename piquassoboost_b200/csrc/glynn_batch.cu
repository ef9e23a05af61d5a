// glynn_batch.cu -- launch unit of the batched small-permanent kernel and the pmf combine kernel.
#include "glynn_batch.cuh"

#include "batch_host.h"

namespace pqb200 {

static_assert(sizeof(BatchJob) == sizeof(BatchJobHost), "BatchJob host/device layout");
static_assert(kBatchMaxRows == kBatchRowsMax && kBatchMaxLen == (int)kBatchLenMax &&
                  kBatchThreads == kBatchCtaThreads, "batch constants");

cudaError_t batch_init_device(void) {
    static double h[kBinomDim][kBinomDim];
    for (int r = 0; r < kBinomDim; r++)
        for (int k = 0; k < kBinomDim; k++) {
            double c = 0.0;
            if (k <= r) {
                c = 1.0;
                for (int t = 1; t <= k; t++) c = c * (double)(r - k + t) / (double)t;  // exact: every prefix is an integer
            }
            h[r][k] = c;
        }
    return cudaMemcpyToSymbol(c_binom, h, sizeof(h));
}

int batch_bucket(int ncols) {
    const int buckets[] = {4, 8, 12, 16, 20, 24, 28, 32};
    for (int b : buckets)
        if (ncols <= b) return b;
    return -1;
}

cudaError_t batch_launch(int ncols_max, const void* d_u, const void* d_jobs, int njobs, void* d_out,
                         cudaStream_t stream) {
    if (njobs <= 0) return cudaSuccess;
    const double2* U = (const double2*)d_u;
    const BatchJob* jobs = (const BatchJob*)d_jobs;
    double2* out = (double2*)d_out;
    switch (ncols_max) {
#define PQB_BATCH_CASE(NM) \
    case NM: glynn_batch_kernel<NM><<<njobs, kBatchThreads, 0, stream>>>(U, jobs, out); break;
        PQB_BATCH_CASE(4)
        PQB_BATCH_CASE(8)
        PQB_BATCH_CASE(12)
        PQB_BATCH_CASE(16)
        PQB_BATCH_CASE(20)
        PQB_BATCH_CASE(24)
        PQB_BATCH_CASE(28)
        PQB_BATCH_CASE(32)
#undef PQB_BATCH_CASE
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t pmf_combine_launch(const void* d_u, int ustride, int m, const void* d_addend,
                               const double* d_weight, double* d_pmf, int shots, cudaStream_t stream) {
    if (shots <= 0) return cudaSuccess;
    pmf_combine_kernel<<<shots, 64, sizeof(double) * (size_t)(m + 1), stream>>>(
        (const double2*)d_u, ustride, m, (const double2*)d_addend, d_weight, d_pmf);
    return cudaGetLastError();
}

}  // namespace pqb200
