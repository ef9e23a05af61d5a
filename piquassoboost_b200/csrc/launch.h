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
    int L;
    unsigned long long chunk_begin, chunk_end;
    int grid;
    double* partials;
    unsigned int* counter;
    double* out;
    cudaStream_t stream;
};

typedef cudaError_t (*plain_launch_fn)(const PlainLaunchArgs&);
typedef int (*plain_occupancy_fn)(void);  // resident CTAs per SM on the current device (<=0: error)

constexpr int kPlainTableSize = 64;
constexpr int kPlainThreads = 256;  // CTA size of glynn_plain_kernel

}  // namespace pqb200
