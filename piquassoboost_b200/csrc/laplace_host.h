// laplace_host.h -- host-visible interface of the launch unit laplace_sampler.cu (the on-device
// photon loop of the Clifford-Clifford B sampler, see laplace_sampler.cuh).
#pragma once
#include <cuda_runtime.h>

namespace pqb200 {

constexpr int kLapPhotonsMax = 40;  // photons per shot
constexpr int kLapModesMax = 255;   // modes are indexed with one byte

struct LapLaunch {
    const void* d_u;  // device, m x m complex128, row stride `ustride`
    int ustride, m, shots, nmax;
    const int* d_nphot;
    const unsigned char* d_ins;
    const double* d_rnd;
    long long* d_samples;
    const unsigned char* d_sample_in;
    const unsigned char* d_cur_in;
    double* d_pmf;
    unsigned long long* d_gray_terms;
    int mode;  // 0: whole shots (fill_r_sample), 1: one pmf layer
};

cudaError_t lap_init_device(void);  // binomial table of the current device
cudaError_t lap_launch(const LapLaunch& l, cudaStream_t stream);

}  // namespace pqb200
