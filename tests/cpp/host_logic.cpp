// Test-only C shim around the HOST headers of libpqb200 (no CUDA involved): lets the CPU test-suite
// check the matrix normalisation and the repeated-row planning (step table, base vectors, binomials)
// by emulating the kernels' sweeps in numpy against the oracle.  Built by tests/test_host_logic_cpu.py.
#include <cstring>
#include <string>

#include "../../piquassoboost_b200/csrc/plain_host.h"
#include "../../piquassoboost_b200/csrc/batch_host.h"
#include "../../piquassoboost_b200/csrc/laplace_host.h"
#include "../../piquassoboost_b200/csrc/repeated_host.h"

using namespace pqb200;

extern "C" {

int hl_plain_cols(int rows, int cols) { return plain_kernel_cols(rows, cols); }

void hl_plain_normalise(const double* a, int rows, int cols, size_t stride, int N, double* a2,
                        double* phase) {
    plain_normalise(a, rows, cols, stride, N, a2, phase);
}

// Prepares + normalises + plans one repeated permanent.  Returns <0 on error, 1 if trivial (value in
// trivial_value), 2 if all multiplicities are 1 (plain route), 0 otherwise.  Caller-allocated
// outputs: mat (R x C complex, columns NOT expanded), r (R-1 ints), cm (C ints), table_w / table_row /
// table_flags (len entries), base (2 x C complex), radix, weight (R-1 entries), binom ((R-1) x 64),
// misc = {R, C, q, len, nchunks, lowpar0, lowpar1, photons}.
int hl_rep_plan(const double* a, size_t rows, size_t cols, size_t stride, const int64_t* row_mult,
                const int64_t* col_mult, int max_len, int normalise, double* trivial_value, double* mat,
                int* r, int* cm, double* table_w, int* table_row, int* table_flags, double* base, int* radix,
                unsigned long long* weight, double* binom, double* phase, long long* misc) {
    RepPrepared pr;
    std::string err;
    const RepLimits lim = {kPowMaxDigits, kPowMaxC, 254, kPowBufBytes, 32, 45};
    int rc = rep_prepare(a, rows, cols, stride, row_mult, col_mult, lim, pr, err);
    if (rc) return -rc;
    if (pr.trivial) {
        trivial_value[0] = pr.value[0];
        trivial_value[1] = pr.value[1];
        return 1;
    }
    misc[0] = pr.R;
    misc[1] = pr.C;
    misc[7] = pr.photons;
    for (int j = 0; j < pr.C; j++) cm[j] = pr.cm[j];
    if (pr.plain_route) {
        memcpy(mat, pr.mat.data(), sizeof(double) * pr.mat.size());
        return 2;
    }
    if (normalise) rep_normalise(pr);
    RepPlan pl;
    rep_plan(pr, 148 * 256, 1, (unsigned int)max_len, 64, pl);
    memcpy(mat, pr.mat.data(), sizeof(double) * pr.mat.size());
    for (int d = 0; d < pr.R - 1; d++) {
        r[d] = pr.r[d];
        radix[d] = pl.radix[d];
        weight[d] = pl.weight[d];
    }
    for (unsigned int t = 0; t < pl.len; t++) {
        table_w[t] = pl.table[t].w;
        table_row[t] = pl.table[t].row;
        table_flags[t] = (int)pl.table[t].flags;
    }
    memcpy(base, pl.base.data(), sizeof(double) * pl.base.size());
    memcpy(binom, pl.binom.data(), sizeof(double) * pl.binom.size());
    memcpy(phase, pr.phase, sizeof(double) * 4);
    misc[2] = pl.q;
    misc[3] = pl.len;
    misc[4] = (long long)pl.nchunks;
    misc[5] = pl.low_parity[0];
    misc[6] = pl.low_parity[1];
    return 0;
}

// launch policy of the B sampler (laplace_host.h): CTAs per shot and photon steps left to the cost-ordered
// second launch, for `shots` shots of `nmax` photons on a device with `sm_count` SMs
int hl_lap_cluster_for(size_t shots, int sm_count) { return lap_cluster_for(shots, sm_count); }
int hl_lap_tail_steps(size_t shots, int nmax, int sm_count) { return lap_tail_steps(shots, nmax, sm_count); }

// One job of the batch path (batch_host.h: index lists only).  Returns make_batch_job's code (0 ok / 2 / 3);
// head = {kind, R, N, q, len, nchunks, transpose}, gidx[R], radix[R-1], pidx[N].
int hl_batch_job(size_t rows, size_t cols, const int64_t* row_mult, const int64_t* col_mult, long long* head,
                 int* gidx, int* radix, int* pidx) {
    BatchJobHost job;
    memset(&job, 0, sizeof(job));
    BatchJobKind kind = kJobKernel;
    std::string err;
    const int rc = make_batch_job(rows, cols, row_mult, col_mult, job, kind, err);
    head[0] = (long long)kind;
    if (rc || kind == kJobTrivialOne || kind == kJobTooBig) return rc;
    head[1] = job.R;
    head[2] = job.N;
    head[3] = job.q;
    head[4] = job.len;
    head[5] = job.nchunks;
    head[6] = job.transpose;
    for (int i = 0; i < job.R; i++) gidx[i] = job.gidx[i];
    for (int i = 0; i + 1 < job.R; i++) radix[i] = job.radix[i];
    for (int i = 0; i < job.N; i++) pidx[i] = job.pidx[i];
    return 0;
}
}
