// laplace_sampler.cu -- launch unit of laplace_sampler_kernel.
#include "laplace_sampler.cuh"

#include "laplace_host.h"

#include <stdio.h>
#include <stdlib.h>

namespace pqb200 {

static_assert(kLapMaxPhotons == kLapPhotonsMax, "laplace constants");

cudaError_t lap_init_device(void) {
    static double h[kLapBinomDim][kLapBinomDim];
    for (int r = 0; r < kLapBinomDim; r++)
        for (int k = 0; k < kLapBinomDim; k++) {
            double c = 0.0;
            if (k <= r) {
                c = 1.0;
                for (int t = 1; t <= k; t++) c = c * (double)(r - k + t) / (double)t;  // every prefix is an integer
            }
            h[r][k] = c;
        }
    return cudaMemcpyToSymbol(c_lap_binom, h, sizeof(h));
}

#ifdef PQB_LAP_PROFILE
static unsigned long long* g_prof = nullptr;
static void lap_prof_report(cudaStream_t stream) {
    unsigned long long h[16];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, g_prof, sizeof(h), cudaMemcpyDeviceToHost);
    const char* names[9] = {"job(warp0)", "gather", "table", "base", "sweep K<8", "sweep K<12", "sweep K<16", "sweep K>=16", "amp+draw"};
    double tot = 0;
    for (int i = 0; i < 9; i++) tot += (double)h[i];
    fprintf(stderr, "[lap profile]");
    for (int i = 0; i < 9; i++) fprintf(stderr, " %s %.1f%%", names[i], 100.0 * (double)h[i] / tot);
    fprintf(stderr, " | total %.3g cycles\n", tot);
    cudaMemset(g_prof, 0, sizeof(h));
}
#endif

cudaError_t lap_order_by_cost(const float* d_cost, int* d_order, int shots, int base, cudaStream_t stream) {
    if (shots <= 0) return cudaSuccess;
    lap_rank_kernel<<<(shots + kLapRankThreads - 1) / kLapRankThreads, kLapRankThreads, 0, stream>>>(d_cost, d_order, shots, base);
    return cudaGetLastError();
}

cudaError_t lap_launch(const LapLaunch& l, cudaStream_t stream) {
    if (l.shots <= 0) return cudaSuccess;
    LapArgs a;
    a.U = (const double2*)l.d_u;
    a.ustride = l.ustride;
    a.m = l.m;
    a.shots = l.shots;
    a.nmax = l.nmax;
    a.nphot = l.d_nphot;
    a.ins = l.d_ins;
    a.rnd = l.d_rnd;
    a.samples_out = l.d_samples;
    a.sample_in = l.d_sample_in;
    a.cur_in = l.d_cur_in;
    a.pmf_out = l.d_pmf;
    a.gray_terms = l.d_gray_terms;
    a.order = l.d_order;
    a.prof = l.d_prof;
    a.step_begin = l.step_begin;
    a.step_end = l.step_end;
    a.state = l.d_state;
    a.cost = l.d_cost;
    a.chunks_per_owner = l.chunks_per_owner > 0 ? l.chunks_per_owner : 4;
    if (const char* f = getenv("PQB200_LAP_CHUNKS")) {  // tuning aid
        const int v = atoi(f);
        if (v >= 1 && v <= 64) a.chunks_per_owner = v;
    }
    a.mode = l.mode;
#ifdef PQB_LAP_PROFILE
    if (!g_prof) {
        cudaMalloc(&g_prof, 16 * sizeof(unsigned long long));
        cudaMemset(g_prof, 0, 16 * sizeof(unsigned long long));
    }
    a.prof = g_prof;
#endif
    const int cs = (l.cluster >= 1 && l.cluster <= kLapMaxCluster) ? l.cluster : 1;
    static_assert(kLapMaxCluster == 8, "cluster sizes");
    if (cs == 1) {  // no cluster: the plain launch
        laplace_sampler_kernel<<<l.shots, kLapThreads, 0, stream>>>(a);
#ifdef PQB_LAP_PROFILE
        lap_prof_report(stream);
#endif
        return cudaGetLastError();
    }
#ifdef PQB_LAP_PROFILE
    struct Report { cudaStream_t s; ~Report() { lap_prof_report(s); } } report__{stream};
#endif
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)l.shots * (unsigned)cs, 1, 1);
    cfg.blockDim = dim3(kLapThreads, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, laplace_sampler_kernel, a);
}

}  // namespace pqb200
