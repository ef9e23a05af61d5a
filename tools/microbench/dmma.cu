#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>  // 0 = DFMA only, 1 = DMMA only, 2 = both
__global__ void __launch_bounds__(256, 1) k(double* out, int iters, double a, double b) {
    double f[8], c[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { f[i] = threadIdx.x * 1e-3 + i; c[i] = i * 0.5 + threadIdx.x; }
    double x = a + threadIdx.x * 1e-9, y = b;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            if (MODE != 1) {
#pragma unroll
                for (int i = 0; i < 8; i++) f[i] = fma(f[i], x, y);
            }
            if (MODE != 0) {
#pragma unroll
                for (int i = 0; i < 8; i += 2) dmma884(c[i], c[i + 1], x, y);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += f[i] + c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
float run(double* d, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, 256>>>(d, 100, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<148, 256>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    double* d; cudaMalloc(&d, 148 * 256 * 8);
    int iters = 100000;
    float t0 = run<0>(d, iters), t1 = run<1>(d, iters), t2 = run<2>(d, iters);
    double nf = 148.0 * 256 * iters * 32;            // DFMA thread-instr
    double nm = 148.0 * 8 * iters * 16;              // DMMA warp-instr (m8n8k4 = 512 flop)
    printf("DFMA only: %.3f ms  %.2f T thread-instr/s (%.2f TFLOP/s)\n", t0, nf / t0 * 1e-9, 2 * nf / t0 * 1e-9);
    printf("DMMA only: %.3f ms  %.3f T warp-instr/s (%.2f TFLOP/s)\n", t1, nm / t1 * 1e-9, 512 * nm / t1 * 1e-9);
    printf("both     : %.3f ms  (sum of separate = %.3f ms; max = %.3f ms)\n", t2, t0 + t1, t0 > t1 ? t0 : t1);
    return 0;
}
