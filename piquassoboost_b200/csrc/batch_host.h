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

// One side (rows or columns) of a job: indices with non-zero multiplicity, and the size of the Gray
// space if this side indexes it: prod(mult+1) * mn/(mn+1), the anchor taking one copy of a
// minimum-multiplicity row.
struct BatchSide {
    int n = 0;                 // entries with non-zero multiplicity
    int64_t total = 0;         // sum of multiplicities
    double space = 1.0;
    int amin = 0;              // position (in idx) of the first minimum-multiplicity entry
    bool bad = false;          // negative multiplicity
    unsigned char idx[256];
    unsigned char mult[256];   // clamped to 255; larger values make the job kJobTooBig via `total`
};

inline void analyze_side(size_t len, const int64_t* mult, BatchSide& sd) {
    sd.n = 0;
    sd.total = 0;
    sd.space = 1.0;
    sd.amin = 0;
    sd.bad = false;
    int64_t mn = INT64_MAX;
    for (size_t i = 0; i < len && i < 256; i++) {
        const int64_t v = mult[i];
        if (v < 0) sd.bad = true;
        if (v <= 0) continue;
        sd.idx[sd.n] = (unsigned char)i;
        sd.mult[sd.n] = (unsigned char)std::min<int64_t>(v, 255);
        if (v < mn) { mn = v; sd.amin = sd.n; }
        sd.space *= (double)(v + 1);
        sd.total += v;
        sd.n++;
    }
    if (sd.n) sd.space *= (double)mn / (double)(mn + 1);
}

// Build the job for perm(U[rows with row_mult>0, cols with col_mult>0]) with multiplicities from
// the two analysed sides.  Returns 0 ok / 2 bad argument / 3 photon mismatch.
inline int make_batch_job_sides(const BatchSide& rs, const BatchSide& cs, BatchJobHost& job,
                                BatchJobKind& kind, const char** err) {
    if (rs.bad || cs.bad) { *err = "negative multiplicity"; return 2; }
    if (rs.total != cs.total) {
        *err = "BBFGPermanentCalculatorRepeated_Tasks::calculate:  Number of input and output states should be equal";
        return 3;
    }
    if (rs.total == 0) { kind = kJobTrivialOne; return 0; }
    if (rs.total > kBatchColsMax) { kind = kJobTooBig; return 0; }
    const bool transpose = cs.space < rs.space;
    const BatchSide& g = transpose ? cs : rs;  // Gray side
    const BatchSide& p = transpose ? rs : cs;  // product side
    if (g.space > kBatchMaxSteps) { kind = kJobTooBig; return 0; }
    job.transpose = transpose ? 1 : 0;
    int R = 0;
    job.gidx[R++] = g.idx[g.amin];
    for (int k = 0; k < g.n; k++) {
        const int mlt = (int)g.mult[k] - (k == g.amin ? 1 : 0);
        if (mlt > 0) {
            if (R >= kBatchRowsMax) { kind = kJobTooBig; return 0; }
            job.radix[R - 1] = (unsigned char)(mlt + 1);
            job.gidx[R++] = g.idx[k];
        }
    }
    job.R = R;
    int N = 0;
    for (int k = 0; k < p.n; k++)
        for (int rep = 0; rep < (int)p.mult[k]; rep++) job.pidx[N++] = p.idx[k];
    job.N = N;
    if (R == 1) { kind = kJobSingleRow; return 0; }
    // low/high split: aim at >= 4 chunks per thread of the CTA when the space allows it
    double total = 1.0;
    for (int d = 0; d < R - 1; d++) total *= job.radix[d];
    const double want_len = std::max(1.0, std::min((double)kBatchLenMax, total / (4.0 * kBatchCtaThreads)));
    unsigned int len = 1;
    int q = 0;
    while (q < R - 1 && (double)len * job.radix[q] <= want_len) { len *= job.radix[q]; q++; }
    job.q = q;
    job.len = len;
    job.nchunks = (unsigned int)(total / (double)len + 0.5);
    kind = kJobKernel;
    return 0;
}

inline int make_batch_job(size_t rows, size_t cols, const int64_t* row_mult, const int64_t* col_mult,
                          BatchJobHost& job, BatchJobKind& kind, std::string& err) {
    if (rows > 255 || cols > 255) {
        int64_t sr = 0, sc = 0;
        for (size_t i = 0; i < rows; i++) sr += row_mult[i];
        for (size_t j = 0; j < cols; j++) sc += col_mult[j];
        if (sr != sc) { err = "BBFGPermanentCalculatorRepeated_Tasks::calculate:  Number of input and output states should be equal"; return 3; }
        kind = (rows == 0 || cols == 0 || sr == 0) ? kJobTrivialOne : kJobTooBig;
        return 0;
    }
    BatchSide rs, cs;
    analyze_side(rows, row_mult, rs);
    analyze_side(cols, col_mult, cs);
    const char* e = nullptr;
    const int rc = make_batch_job_sides(rs, cs, job, kind, &e);
    if (rc) err = e;
    return rc;
}

// ---- implemented in glynn_batch.cu
cudaError_t batch_init_device(void);  // fills the constant binomial table on the current device
cudaError_t batch_launch(int ncols_max, const void* d_u, const void* d_jobs, int njobs, void* d_out,
                         cudaStream_t stream);
cudaError_t pmf_combine_launch(const void* d_u, int ustride, int m, const void* d_addend,
                               const double* d_weight, double* d_pmf, int shots, cudaStream_t stream);
int batch_bucket(int ncols);  // smallest instantiated NMAX >= ncols, or -1

}  // namespace pqb200
