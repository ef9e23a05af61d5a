// batch_host.h -- host-side job construction for glynn_batch_kernel (index lists only: no matrix
// data is touched on the host) and the interface of the batch launch unit glynn_batch.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <string>
#include <vector>

namespace pqb200 {

constexpr int kBatchRowsMax = 32;
constexpr int kBatchColsMax = 32;
constexpr unsigned int kBatchLenMax = 512;
constexpr int kBatchCtaThreads = 128;
constexpr double kBatchMaxSteps = 4194304.0;  // jobs with a larger Gray space go to the single-permanent path

// must match BatchJob in glynn_batch.cuh (checked with static_assert in glynn_batch.cu)
struct BatchJobHost {
    int R, N, q;
    unsigned int len, nchunks;
    int transpose, out_slot, ustride;
    unsigned long long uoff;
    unsigned char gidx[32];
    unsigned char radix[32];
    unsigned char pidx[48];
};

enum BatchJobKind { kJobKernel = 0, kJobTrivialOne = 1, kJobSingleRow = 2, kJobTooBig = 3 };

// Build the job for perm(U[rows with row_mult>0, cols with col_mult>0]) with multiplicities.
// Returns 0 ok / 2 bad argument / 3 photon mismatch.  `kind` tells the caller what to do with it.
inline int make_batch_job(size_t rows, size_t cols, const int64_t* row_mult, const int64_t* col_mult,
                          BatchJobHost& job, BatchJobKind& kind, std::string& err) {
    int64_t sr = 0, sc = 0;
    for (size_t i = 0; i < rows; i++) { if (row_mult[i] < 0) { err = "negative row multiplicity"; return 2; } sr += row_mult[i]; }
    for (size_t j = 0; j < cols; j++) { if (col_mult[j] < 0) { err = "negative column multiplicity"; return 2; } sc += col_mult[j]; }
    if (sr != sc) {
        err = "BBFGPermanentCalculatorRepeated_Tasks::calculate:  Number of input and output states should be equal";
        return 3;
    }
    if (rows == 0 || cols == 0 || sr == 0) { kind = kJobTrivialOne; return 0; }
    if (rows > 255 || cols > 255 || sr > kBatchColsMax) { kind = kJobTooBig; return 0; }
    std::vector<int> ri, ci;
    for (size_t i = 0; i < rows; i++) if (row_mult[i] > 0) ri.push_back((int)i);
    for (size_t j = 0; j < cols; j++) if (col_mult[j] > 0) ci.push_back((int)j);
    auto space = [](const std::vector<int>& idx, const int64_t* mult) {
        double s = 1.0;
        int64_t mn = INT64_MAX;
        for (int i : idx) { s *= (double)(mult[i] + 1); mn = std::min(mn, mult[i]); }
        return s * (double)mn / (double)(mn + 1);
    };
    const double srow = space(ri, row_mult), scol = space(ci, col_mult);
    const bool transpose = scol < srow;
    const std::vector<int>& gi = transpose ? ci : ri;
    const std::vector<int>& pi = transpose ? ri : ci;
    const int64_t* gm = transpose ? col_mult : row_mult;
    const int64_t* pm = transpose ? row_mult : col_mult;
    if (std::min(srow, scol) > kBatchMaxSteps) { kind = kJobTooBig; return 0; }
    int amin = 0;
    for (int k = 1; k < (int)gi.size(); k++) if (gm[gi[k]] < gm[gi[amin]]) amin = k;
    job.transpose = transpose ? 1 : 0;
    int R = 0;
    job.gidx[R++] = (unsigned char)gi[amin];
    for (int k = 0; k < (int)gi.size(); k++) {
        const int mlt = (int)gm[gi[k]] - (k == amin ? 1 : 0);
        if (mlt > 0) {
            if (R >= kBatchRowsMax) { kind = kJobTooBig; return 0; }
            job.radix[R - 1] = (unsigned char)(mlt + 1);
            job.gidx[R++] = (unsigned char)gi[k];
        }
    }
    job.R = R;
    int N = 0;
    for (int p : pi) for (int64_t rep = 0; rep < pm[p]; rep++) job.pidx[N++] = (unsigned char)p;
    job.N = N;
    if (R == 1) { kind = kJobSingleRow; return 0; }
    // low/high split: keep at least ~one chunk per thread of the CTA when the space allows it
    double total = 1.0;
    for (int d = 0; d < R - 1; d++) total *= job.radix[d];
    const double want_len = std::max(1.0, std::min((double)kBatchLenMax, total / kBatchCtaThreads));
    unsigned int len = 1;
    int q = 0;
    while (q < R - 1 && (double)len * job.radix[q] <= want_len) { len *= job.radix[q]; q++; }
    job.q = q;
    job.len = len;
    job.nchunks = (unsigned int)(total / (double)len + 0.5);
    kind = kJobKernel;
    return 0;
}

// ---- implemented in glynn_batch.cu
cudaError_t batch_init_device(void);  // fills the constant binomial table on the current device
cudaError_t batch_launch(int ncols_max, const void* d_u, const void* d_jobs, int njobs, void* d_out,
                         cudaStream_t stream);
cudaError_t pmf_combine_launch(const void* d_u, int ustride, int m, const void* d_addend,
                               const double* d_weight, double* d_pmf, int shots, cudaStream_t stream);
int batch_bucket(int ncols);  // smallest instantiated NMAX >= ncols, or -1

}  // namespace pqb200
