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

struct RepLaunchArgs {
    const double* mat;  // host, interleaved complex, R x N (row 0 = anchor, columns physically repeated)
    int R, N;
    int q;
    unsigned int len;
    unsigned long long chunk_begin, chunk_end;
    int grid;
    const unsigned char* radix;        // [R-1]
    const unsigned long long* weight;  // [R-1]
    const void* table;                 // HOST, RepStepHost[len] (copied into the kernel parameters)
    const void* base;                  // device, double2[2][N]
    const double* binom;               // device, [R-1][kRepBinomStride]
    int low_parity[2];
    double phase[4];                   // dd complex correction factor of the normalisation
    double* partials;
    unsigned int* counter;
    double* out;
    cudaStream_t stream;
};

typedef cudaError_t (*rep_launch_fn)(const RepLaunchArgs&);

constexpr int kRepBinomStride = 45;  // kRepMaxN + 1
constexpr int kPlainTableSize = 64;
constexpr int kPlainThreads = 256;  // CTA size of glynn_plain_kernel

}  // namespace pqb200
