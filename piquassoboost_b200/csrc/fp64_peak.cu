// fp64_peak.cu -- measures the FP64 pipe's issue rate on this device (the roofline denominator of
// the permanent kernels; MEASURED_PEAKS.json has no FP64 entry).  Independent dependency chains of
// DFMA / DADD / DMUL, 8 per thread, 2 CTAs of 256 threads per SM, timed with CUDA events.
#include <cuda_runtime.h>

#include "../../include/pqb200.h"

namespace {

template <int OP>
__global__ void __launch_bounds__(256) fp64_chain_kernel(double* out, int iters, double a, double b) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (OP == 0) x[i] = fma(x[i], a, b);
                if (OP == 1) x[i] = x[i] + a;
                if (OP == 2) x[i] = x[i] * a;
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    if (s == 12345.678) out[0] = s;  // never true: keeps the chains alive
}

template <int OP>
cudaError_t run(int sms, int iters, double* seconds) {
    double* d = nullptr;
    cudaError_t e = cudaMalloc(&d, 8);
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    fp64_chain_kernel<OP><<<sms * 2, 256>>>(d, 200, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    fp64_chain_kernel<OP><<<sms * 2, 256>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    e = cudaDeviceSynchronize();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    *seconds = ms * 1e-3;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return e;
}

}  // namespace

extern "C" int pqb200_fp64_peak(int device, int op, double duration_hint_s, double* thread_instr_per_s) {
    if (!thread_instr_per_s || op < 0 || op > 2) return PQB200_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return PQB200_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PQB200_ERR_CUDA;
    const int sms = prop.multiProcessorCount;
    // ~17e12 thread-instr/s on B200: pick iters for the requested duration
    double want = duration_hint_s > 0 ? duration_hint_s : 0.02;
    int iters = (int)(want * 17e12 / ((double)sms * 2 * 256 * 64));
    if (iters < 1000) iters = 1000;
    double sec = 0;
    cudaError_t e = op == 0 ? run<0>(sms, iters, &sec) : op == 1 ? run<1>(sms, iters, &sec) : run<2>(sms, iters, &sec);
    if (e != cudaSuccess || sec <= 0) return PQB200_ERR_CUDA;
    *thread_instr_per_s = (double)sms * 2 * 256 * (double)iters * 64.0 / sec;
    return PQB200_OK;
}
