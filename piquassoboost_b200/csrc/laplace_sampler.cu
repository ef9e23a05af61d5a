// laplace_sampler.cu -- launch unit of laplace_sampler_kernel.
#include "laplace_sampler.cuh"

#include "laplace_host.h"

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
    a.chunks_per_owner = l.chunks_per_owner > 0 ? l.chunks_per_owner : 8;
    if (const char* f = getenv("PQB200_LAP_CHUNKS")) {  // tuning aid
        const int v = atoi(f);
        if (v >= 1 && v <= 64) a.chunks_per_owner = v;
    }
    a.mode = l.mode;
    const int cs = (l.cluster == 2 || l.cluster == 4 || l.cluster == 8) ? l.cluster : 1;
    static_assert(kLapMaxCluster == 8, "cluster sizes");
    if (cs == 1) {  // no cluster: the plain launch
        laplace_sampler_kernel<<<l.shots, kLapThreads, 0, stream>>>(a);
        return cudaGetLastError();
    }
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
