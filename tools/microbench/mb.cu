// tools/microbench/mb.cu -- round-1 design experiments run on the B200 box (not part of the product).
//   1. FP64 pipe peak (DFMA / DADD / DMUL chains) -> lanes per SM per clock, sustained clock
//   2. candidate Gray-kernel variants timed on the same n=40-shaped workload (restricted chunk range)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <complex>
#include <random>
#include <cuda_runtime.h>
#include "variant_static.cuh"
#include "variant_rolled.cuh"
#include "variant_uniform.cuh"
using namespace pqb200;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

template <int ILP, int OP>
__global__ void __launch_bounds__(256) fp64_peak(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                if (OP == 0) x[i] = fma(x[i], a, b);
                if (OP == 1) x[i] = x[i] + a;
                if (OP == 2) x[i] = x[i] * a;
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    if (s == 12345.678) out[0] = s;
}

template <int ILP, int OP>
void run_peak(const char* name, int blocks_per_sm, int threads) {
    double* d; CK(cudaMalloc(&d, 8));
    int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int grid = 148 * blocks_per_sm;
    fp64_peak<ILP, OP><<<grid, threads>>>(d, 100, 1.0000001, 1e-9);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    fp64_peak<ILP, OP><<<grid, threads>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double instr = (double)grid * threads * iters * 8.0 * ILP;
    printf("PEAK %-6s ilp=%d blocks/SM=%d threads=%d : %.3f ms  %.3f T thread-instr/s  (=%.2f lanes/SM/clk @1.965GHz)  %s TFLOP/s=%.2f\n",
           name, ILP, blocks_per_sm, threads, ms, instr / ms / 1e9, instr / (ms * 1e-3) / 148 / 1.965e9,
           OP == 0 ? "fma" : "", (OP == 0 ? 2 : 1) * instr / ms / 1e9);
    cudaFree(d);
}

typedef std::complex<long double> cld;
static cld cpu_glynn(const std::vector<std::complex<double>>& A, int n) {
    std::vector<cld> cs(n);
    for (int j = 0; j < n; j++) { cs[j] = 0; for (int i = 0; i < n; i++) cs[j] += (cld)A[i * n + j]; }
    cld sum = 0; unsigned long long T = 1ull << (n - 1); unsigned long long gray = 0; int sign = 1;
    for (unsigned long long g = 0; g < T; g++) {
        if (g) { unsigned long long ng = g ^ (g >> 1); unsigned long long diff = ng ^ gray; int b = __builtin_ctzll(diff); int row = n - 1 - b;
            long double s = (ng >> b) & 1 ? -2.0L : 2.0L; for (int j = 0; j < n; j++) cs[j] += s * (cld)A[row * n + j]; gray = ng; sign = -sign; }
        cld p = 1; for (int j = 0; j < n; j++) p *= cs[j];
        sum += (long double)sign * p;
    }
    return sum / (long double)T;
}

struct Bufs { double* partials; unsigned int* counter; double* out; double2* dA; };

template <int N>
void run_variants(int rows, int L, unsigned long long nchunks_to_run, bool check) {
    const int n = rows;
    std::mt19937_64 rng(n); std::normal_distribution<double> nd(0, 1);
    std::vector<std::complex<double>> A(n * N);
    for (auto& z : A) z = std::complex<double>(nd(rng), nd(rng)) * std::sqrt(0.5 / n);
    Bufs b; CK(cudaMalloc(&b.partials, 4 * 8 * 4096)); CK(cudaMalloc(&b.counter, 4)); CK(cudaMemset(b.counter, 0, 4)); CK(cudaMalloc(&b.out, 32));
    CK(cudaMalloc(&b.dA, 16 * n * N)); CK(cudaMemcpy(b.dA, A.data(), 16 * n * N, cudaMemcpyHostToDevice));
    unsigned long long total_chunks = 1ull << (n - 1 - L);
    if (nchunks_to_run > total_chunks) nchunks_to_run = total_chunks;
    cld ref = 0; if (check) ref = cpu_glynn(A, n);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double steps = (double)nchunks_to_run * (double)(1ull << L);
    int grid = (int)std::min<unsigned long long>(148, (nchunks_to_run + 255) / 256);
    auto report = [&](const char* name, float ms) {
        double h[4]; CK(cudaMemcpy(h, b.out, 32, cudaMemcpyDeviceToHost));
        double sc = std::ldexp(1.0, -(n - 1));
        double re = (h[0] + h[1]) * sc, im = (h[2] + h[3]) * sc;
        double err = check ? (double)(std::abs(cld(re, im) - ref) / std::abs(ref)) : -1;
        printf("VAR %-10s N=%d rows=%d L=%d chunks=%llu grid=%d : %.3f ms  %.4e steps/s  FP64 %.2f TFLOP/s (8N/step)  val=(%.15e,%.15e) relerr=%.3e\n",
               name, N, rows, L, nchunks_to_run, grid, ms, steps / (ms * 1e-3), steps * 8 * N / ms / 1e9, re, im, err);
    };
    // --- variant U
    {
        static UParams<N, N> p;
        for (int i = 0; i < n; i++) for (int j = 0; j < N; j++) p.A2[i][j] = make_double2(2 * A[i * N + j].real(), 2 * A[i * N + j].imag());
        p.rows = n; p.L = L; p.chunk_begin = 0; p.chunk_end = nchunks_to_run;
        p.iters = (unsigned)((nchunks_to_run + (unsigned long long)grid * 256 - 1) / ((unsigned long long)grid * 256));
        p.partials = b.partials; p.counter = b.counter; p.out = b.out;
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            glynn_u_kernel<N, N><<<grid, 256>>>(p);
            cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) report("uniform", ms);
        }
    }
    // --- rolled LDS / LDC
    {
        static RolledParams<N, N> p;
        for (int i = 0; i < n; i++) for (int j = 0; j < N; j++) p.A2c[i][j] = make_double2(2 * A[i * N + j].real(), 2 * A[i * N + j].imag());
        p.A = b.dA; p.rows = n; p.L = L; p.chunk_begin = 0; p.chunk_end = nchunks_to_run;
        p.partials = b.partials; p.counter = b.counter; p.out = b.out;
        CK(cudaFuncSetAttribute(glynn_rolled_kernel<N, 0, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * n * N));
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            glynn_rolled_kernel<N, 0, N><<<grid, 256, 16 * n * N>>>(p);
            cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) report("rolledLDS", ms);
        }
    }
    // --- static S=2 + LDS runtime rows
    {
        constexpr int S = 2;
        static PlainParams<N, S> p;
        for (int t = 1; t < (1 << S); t++) {
            int bb = __builtin_ctz(t); bool set = ((t >> bb) ^ (t >> (bb + 1))) & 1; int row = n - 1 - bb;
            for (int j = 0; j < N; j++) p.low[t - 1][j] = make_double2((set ? -2 : 2) * A[row * N + j].real(), (set ? -2 : 2) * A[row * N + j].imag());
        }
        p.A = b.dA; p.rows = n; p.L = L; p.chunk_begin = 0; p.chunk_end = nchunks_to_run;
        p.partials = b.partials; p.counter = b.counter; p.out = b.out;
        CK(cudaFuncSetAttribute(glynn_plain_kernel<N, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * n * N));
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            glynn_plain_kernel<N, S><<<grid, 256, 16 * n * N>>>(p);
            cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) report("staticS2", ms);
        }
    }
    cudaFree(b.partials); cudaFree(b.counter); cudaFree(b.out); cudaFree(b.dA);
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s SMs=%d clock=%d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
    run_peak<8, 0>("DFMA", 1, 256); run_peak<8, 0>("DFMA", 2, 256); run_peak<8, 0>("DFMA", 4, 256); run_peak<8, 0>("DFMA", 8, 256);
    run_peak<4, 0>("DFMA", 1, 256); run_peak<2, 0>("DFMA", 1, 256); run_peak<1, 0>("DFMA", 1, 256); run_peak<1, 0>("DFMA", 1, 128);
    run_peak<8, 1>("DADD", 2, 256); run_peak<8, 2>("DMUL", 2, 256);
    // correctness at small n
    run_variants<16>(16, 6, 1ull << 40, true);
    run_variants<24>(24, 10, 1ull << 40, true);
    // n=40-shaped timing: 2^18 chunks of 2^12 steps = 2^30 steps
    run_variants<40>(40, 12, 1ull << 18, false);
    run_variants<40>(40, 12, 1ull << 20, false);
    run_variants<32>(32, 12, 1ull << 19, false);
    run_variants<30>(30, 12, 1ull << 17, false);
    return 0;
}
