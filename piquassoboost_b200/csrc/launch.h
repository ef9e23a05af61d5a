// launch.h -- host-side launch descriptors shared by pqb200_capi.cu and the per-N kernel
// instantiation units (the kernels are templated on the column count N; each translation unit
// instantiates a range of N and registers its launchers in a table).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace pqb200 {

struct PlainLaunchArgs {
    const double* a;  // host, interleaved complex, rows x cols, `stride` complex elements per row
    size_t rows, cols, stride;
    int L, K;  // log2 chunk / sub-chunk length
    unsigned long long chunk_begin, chunk_end;
    int grid;
    double* partials;
    unsigned int* counter;
    double* out;
    cudaStream_t stream;
};

typedef cudaError_t (*plain_launch_fn)(const PlainLaunchArgs&);
typedef int (*plain_occupancy_fn)(void);  // resident CTAs per SM on the current device (<=0: error)

// ---- power-form repeated kernel (glynn_reppow.cuh)
constexpr int kPowMaxC = 48;        // distinct product-side columns
constexpr int kPowMaxDigits = 44;   // Gray-side rows besides the anchor
constexpr int kPowBufBytes = 32128; // matrix + step tables inside the kernel-parameter bank
constexpr int kPowLenCap = 4096;
constexpr int kPowTableSize = 64;
// the double-double walk (dd_repeated.cu) keeps its column sums in local memory: wider limits, which also
// serve shapes beyond the fast kernel's
constexpr int kDDRepCols = 128;
constexpr int kDDRepDigits = 64;

#ifdef __CUDACC__
#define PQB_HD __host__ __device__
#else
#define PQB_HD
#endif
PQB_HD constexpr int pow_row_stride(int c) { return c % 16 == 0 ? c + 1 : c; }
// resident CTAs per SM the kernel's register budget is set for (its __launch_bounds__), also what the
// host plans the chunk decomposition with
#ifndef PQB_POW_MINB
#define PQB_POW_MINB(c) ((c) <= 8 ? 3 : (c) <= 26 ? 2 : 1)
#endif
PQB_HD constexpr int pow_min_blocks(int c) { return PQB_POW_MINB(c); }
// steps per chunk that fit the parameter bank next to an R x C matrix (10 bytes per step)
inline int pow_len_max(int R, int C) {
    const int v = (kPowBufBytes - 16 * R * pow_row_stride(C) - 64) / 10;
    return v > kPowLenCap ? kPowLenCap : v;
}

struct PowLaunchArgs {
    const double* mat;  // host, interleaved complex, R x C (row 0 = anchor), normalised, NOT doubled
    int R, C;
    const unsigned char* cmult;  // [C] column multiplicities (>= 1), in the kernel's column order
    int q;
    unsigned int len;
    unsigned long long chunk_begin, chunk_end;
    int sm_count;                      // the launcher sizes the grid: resident CTAs per SM x sm_count
    const unsigned char* radix;        // [R-1]
    const unsigned long long* weight;  // [R-1]
    const void* table;                 // HOST, RepStepHost[len] (copied into the kernel parameters)
    const void* base;                  // device, double2[2][C]
    const double* binom;               // device, [R-1][binom_stride]
    int binom_stride;
    int low_parity[2];
    double phase[4];
    double* partials;                  // [grid][8]
    unsigned int* counter;
    double* out;                       // 5 doubles
    cudaStream_t stream;
};
typedef cudaError_t (*pow_launch_fn)(const PowLaunchArgs&);

constexpr int kPlainTableSize = 64;
constexpr int kPlainThreads = 256;  // CTA size of glynn_plain_kernel

}  // namespace pqb200
