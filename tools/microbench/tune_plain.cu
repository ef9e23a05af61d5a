// tools/microbench/tune_plain.cu -- per-N tuning sweep of glynn_plain_kernel<N, CH, U> (multiply chains
// x loop form).  ptxas' schedule decides whether the FP64 pipe stays fed, and it is sensitive to N:
// the (CH, U) table in glynn_plain.cuh (default_chains / default_unroll) is the argmin of this sweep.
// Build one binary per N range:  nvcc -DNLO=.. -DNHI=.. ; each prints "TUNE N CH U ms steps/s".
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <complex>
#include <random>
#include <cuda_runtime.h>
#include "../../piquassoboost_b200/csrc/glynn_plain.cuh"
#include "../../piquassoboost_b200/csrc/plain_host.h"
using namespace pqb200;
#ifndef KSUB
#define KSUB 6
#endif
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

template <int N, int CH, int U>
struct Kern { static constexpr auto fn = glynn_plain_kernel<N, CH, U>; };
template <int N, int CH>
struct Kern<N, CH, 2> { static constexpr auto fn = glynn_plain_rolled_kernel<N, CH>; };
template <int N, int CH>
struct Kern<N, CH, 22> { static constexpr auto fn = glynn_plain_rolled_kernel<N, CH, 2>; };  // 2 CTAs/SM

template <int N, int CH, int U>
void run_variant(const std::vector<double>& A, int L, unsigned long long nchunks, double* partials, unsigned int* counter, double* out) {
    static PlainParams<N> p;
    plain_normalise(A.data(), N, N, N, N, reinterpret_cast<double*>(&p.A2[0][0]), p.phase, plain_row_stride(N));
    CK(cudaFuncSetAttribute(Kern<N, CH, U>::fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plain_anchor_bytes<N>()));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, Kern<N, CH, U>::fn, 256, plain_anchor_bytes<N>()));
    int grid = 148 * occ;
    p.fast_row = 0; p.rows = N; p.L = L; p.K = KSUB < L ? KSUB : L; p.chunk_begin = 0; p.chunk_end = nchunks;
    p.iters = (unsigned)((nchunks + (unsigned long long)grid * 256 - 1) / ((unsigned long long)grid * 256));
    p.partials = partials; p.counter = counter; p.out = out;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        Kern<N, CH, U>::fn<<<grid, 256, plain_anchor_bytes<N>()>>>(p);
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep) best = std::min(best, ms);
    }
    double h[4]; CK(cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, Kern<N, CH, U>::fn));
    const double steps = (double)nchunks * (double)(1ull << L);
    printf("TUNE %d %d %d %.3f %.4e regs=%d local=%zu sum=(%.15e,%.15e)\n", N, CH, U, best, steps / (best * 1e-3),
           fa.numRegs, fa.localSizeBytes, h[0] + h[1], h[2] + h[3]);
}

template <int N>
void sweep() {
    const int L = N - 1 < 12 ? N - 1 : 12;
    // ~2^31 steps for big N; everything for small N
    // always 2^31 steps: for small N the chunk range simply revisits the Gray space (throughput test)
    const int lg = 31;
    const unsigned long long nchunks = 1ull << (lg - L);
    std::mt19937_64 rng(N); std::normal_distribution<double> nd(0, 1);
    std::vector<double> A(2 * N * N);
    for (auto& z : A) z = nd(rng) * std::sqrt(0.5 / N);
    double* partials; unsigned int* counter; double* out;
    CK(cudaMalloc(&partials, 4 * 8 * 4096)); CK(cudaMalloc(&counter, 4)); CK(cudaMemset(counter, 0, 4)); CK(cudaMalloc(&out, 32));
#ifdef TWO_CTA
    run_variant<N, 2, 22>(A, L, nchunks, partials, counter, out);
    run_variant<N, 3, 22>(A, L, nchunks, partials, counter, out);
    run_variant<N, 4, 22>(A, L, nchunks, partials, counter, out);
    run_variant<N, 6, 22>(A, L, nchunks, partials, counter, out);
#endif
    run_variant<N, 2, 2>(A, L, nchunks, partials, counter, out);
    run_variant<N, 3, 2>(A, L, nchunks, partials, counter, out);
    run_variant<N, 4, 2>(A, L, nchunks, partials, counter, out);
    run_variant<N, 6, 2>(A, L, nchunks, partials, counter, out);
    run_variant<N, 8, 2>(A, L, nchunks, partials, counter, out);
    run_variant<N, 2, 4>(A, L, nchunks, partials, counter, out);
    run_variant<N, 4, 4>(A, L, nchunks, partials, counter, out);
    run_variant<N, 8, 4>(A, L, nchunks, partials, counter, out);
    run_variant<N, 2, 8>(A, L, nchunks, partials, counter, out);
    run_variant<N, 4, 8>(A, L, nchunks, partials, counter, out);
    run_variant<N, 8, 8>(A, L, nchunks, partials, counter, out);
    run_variant<N, 5, 2>(A, L, nchunks, partials, counter, out);
    run_variant<N, 7, 2>(A, L, nchunks, partials, counter, out);
    run_variant<N, 6, 8>(A, L, nchunks, partials, counter, out);
    run_variant<N, 3, 8>(A, L, nchunks, partials, counter, out);
    cudaFree(partials); cudaFree(counter); cudaFree(out);
}

template <int LO, int HI>
struct Range { static void go() { sweep<LO>(); Range<LO + 1, HI>::go(); } };
template <int HI>
struct Range<HI, HI> { static void go() { sweep<HI>(); } };

int main() { Range<NLO, NHI>::go(); return 0; }
