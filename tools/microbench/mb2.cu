// tools/microbench/mb2.cu -- round-1 tuning sweep of the PRODUCT kernel (glynn_plain_kernel) on an
// n=40-shaped workload: dd-accumulation period DDP and multiply-chain count CH; plus FP64 pipe
// rates with the SM clock measured inside the kernel (clock64 / globaltimer).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <complex>
#include <random>
#include <cuda_runtime.h>
#include "../../piquassoboost_b200/csrc/glynn_plain.cuh"
using namespace pqb200;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

template <int OP, int NREG>
__global__ void __launch_bounds__(256) fp64_rate(double* out, unsigned long long* clk, int iters, double a, double b) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-9 + i;
    double y = a * 0.5, z = b * 2.0;  // register operands
    unsigned long long c0 = clock64(), t0 = gtimer();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (OP == 0) x[i] = NREG == 3 ? fma(x[i], y, z) : fma(x[i], a, b);
                if (OP == 1) x[i] = x[i] + (NREG == 3 ? y : a);
                if (OP == 2) x[i] = x[i] * (NREG == 3 ? y : a);
                if (OP == 3) { if (i & 1) x[i] = fma(x[i], y, z); else x[i] = x[i] * y; }
            }
        }
        if (NREG == 3) { y += 1e-30; z += 1e-30; }
    }
    unsigned long long c1 = clock64(), t1 = gtimer();
    if (blockIdx.x == 0 && threadIdx.x == 0) { clk[0] = c1 - c0; clk[1] = t1 - t0; }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    if (s == 12345.678) out[0] = s;
}

template <int OP, int NREG>
void run_rate(const char* name, int iters) {
    double* d; unsigned long long* clk; CK(cudaMalloc(&d, 8)); CK(cudaMalloc(&clk, 16));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    fp64_rate<OP, NREG><<<296, 256>>>(d, clk, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    fp64_rate<OP, NREG><<<296, 256>>>(d, clk, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[2]; CK(cudaMemcpy(h, clk, 16, cudaMemcpyDeviceToHost));
    double mhz = (double)h[0] / (double)h[1] * 1e3;
    double instr = 296.0 * 256 * iters * 64.0;
    printf("RATE %-14s iters=%d: %.3f ms  %.3f T thread-instr/s  SM clock %.0f MHz  -> %.2f lanes/SM/clk\n", name, iters, ms,
           instr / ms / 1e9, mhz, instr / (ms * 1e-3) / 148 / (mhz * 1e6));
}

template <int N, int CH, int U>
void run_variant(const std::vector<std::complex<double>>& A, int n, int L, unsigned long long nchunks, double* partials, unsigned int* counter, double* out) {
    static PlainParams<N> p;
    for (int i = 0; i < n; i++) for (int j = 0; j < N; j++) p.A2[i][j] = make_double2(2 * A[i * N + j].real(), 2 * A[i * N + j].imag());
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, glynn_plain_kernel<N, CH, U>, 256, 0));
    int grid = 148 * occ;
    p.rows = n; p.L = L; p.chunk_begin = 0; p.chunk_end = nchunks;
    p.iters = (unsigned)((nchunks + (unsigned long long)grid * 256 - 1) / ((unsigned long long)grid * 256));
    p.partials = partials; p.counter = counter; p.out = out;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        glynn_plain_kernel<N, CH, U><<<grid, 256>>>(p);
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep) best = std::min(best, ms);
    }
    double h[4]; CK(cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, glynn_plain_kernel<N, CH, U>));
    const double steps = (double)nchunks * (double)(1ull << L);
    printf("VAR N=%d CH=%d U=%d L=%d regs=%d local=%zu occ=%d: %.3f ms  %.4e steps/s  %.2f TFLOP/s  sum=(%.12e,%.12e)\n", N, CH, U, L,
           fa.numRegs, fa.localSizeBytes, occ, best, steps / (best * 1e-3), steps * 8 * N / best / 1e9, h[0] + h[1], h[2] + h[3]);
}

template <int N>
void sweep(int L, unsigned long long nchunks) {
    const int n = N;
    std::mt19937_64 rng(n); std::normal_distribution<double> nd(0, 1);
    std::vector<std::complex<double>> A(n * N);
    for (auto& z : A) z = std::complex<double>(nd(rng), nd(rng)) * std::sqrt(0.5 / n);
    double* partials; unsigned int* counter; double* out;
    CK(cudaMalloc(&partials, 4 * 8 * 4096)); CK(cudaMalloc(&counter, 4)); CK(cudaMemset(counter, 0, 4)); CK(cudaMalloc(&out, 32));
    run_variant<N, 2, 4>(A, n, L, nchunks, partials, counter, out);
    run_variant<N, 4, 4>(A, n, L, nchunks, partials, counter, out);
    run_variant<N, 8, 4>(A, n, L, nchunks, partials, counter, out);
    run_variant<N, 2, 8>(A, n, L, nchunks, partials, counter, out);
    run_variant<N, 4, 8>(A, n, L, nchunks, partials, counter, out);
    run_variant<N, 8, 8>(A, n, L, nchunks, partials, counter, out);
    cudaFree(partials); cudaFree(counter); cudaFree(out);
}

int main() {
    run_rate<0, 1>("DFMA R,UR,UR", 40000); run_rate<0, 3>("DFMA R,R,R", 40000);
    sweep<40>(12, 1ull << 19);
    sweep<38>(12, 1ull << 19);
    sweep<36>(12, 1ull << 19);
    sweep<34>(12, 1ull << 19);
    sweep<32>(12, 1ull << 19);
    return 0;
}
