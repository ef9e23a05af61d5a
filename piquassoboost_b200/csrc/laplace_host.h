// laplace_host.h -- host-visible interface of the launch unit laplace_sampler.cu (the on-device
// photon loop of the Clifford-Clifford B sampler, see laplace_sampler.cuh).
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

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
    const int* d_order;  // [shots] shot index handled by cluster i (NULL = identity)
    int cluster;         // CTAs per shot: 1 .. 8
    int chunks_per_owner;  // low/high digit split: chunks aimed at per term owner (0 = default 4, measured)
    unsigned long long* d_prof;  // PQB_LAP_PROFILE builds only
    int step_begin, step_end;    // mode 0: photon steps [step_begin, step_end) of every shot (0, nmax = all)
    unsigned char* d_state;      // mode 0: [shots][2][m] hand-over state between launches over step ranges
    float* d_cost;               // mode 0: [shots] cost estimate written at the hand-over
    int mode;  // 0: whole shots (fill_r_sample), 1: one pmf layer
};

// Photon steps left to the second, cost-ordered launch of a wave (0: a single launch).
inline int lap_tail_steps(size_t shots, int nmax, int sm_count) {
    // measured (profiles/sampler_sweep_tail_r2s.log, kernel ms at tail 0 / 3 / 4 / 5): 1000 shots 3.9 / 3.18 / 3.11 /
    // 3.11; 444 shots (every CTA resident at once: the ordered launch still spreads the long shots over the SMs)
    // 2.30 / - / 2.17; 125 shots (clusters) 0.94 / - / 0.98; waves of 5000 shots 56.7 / 56.8 / 56.0 / 56.1
    int tail = 4;
    const size_t slots = (size_t)sm_count * 3;
    if (nmax < 10 || shots < slots || shots > 6 * slots) tail = 0;  // clusters fill the SMs / enough shots to balance
    if (const char* f = getenv("PQB200_LAP_TAIL")) {  // tuning aid
        const int v = atoi(f);
        if (v >= 0 && v < nmax) tail = v;
    }
    return tail;
}

// CTAs per shot: the largest cluster that still fits the device's resident CTA slots in one wave
inline int lap_cluster_for(size_t shots, int sm_count) {
    if (const char* f = getenv("PQB200_LAP_CLUSTER")) {  // tuning aid
        const int v = atoi(f);
        if (v >= 1 && v <= 8) return v;
    }
    // measured (tools/sampler_sweep.py, profiles/sampler_sweep_r2.log): every CTA of a cluster repeats the
    // per-step preparation, so splitting only pays while it fills otherwise idle SMs
    // so: the largest cluster whose CTAs are all resident at once; clusters of 8 synchronise more dearly and
    // only pay with SMs to spare (measured kernel ms for 25 / 50 / 100 / 200 shots of C5 at 1, 2, 4, 8 CTAs
    // per shot: 1.40 0.86 0.58 0.48 / 1.63 0.99 0.69 0.78 / 1.75 1.13 0.83 1.22 / 1.87 1.27 1.36 2.03)
    const size_t slots = (size_t)sm_count * 3;
    if (shots == 0) return 1;
    size_t cs = slots / shots;  // any size up to the portable limit of 8 is a valid cluster: 125 shots get 3
    if (cs > 4) cs = shots * 8 * 3 <= slots * 2 ? 8 : 4;
    return cs < 1 ? 1 : (int)cs;
}

cudaError_t lap_init_device(void);  // binomial table of the current device
cudaError_t lap_launch(const LapLaunch& l, cudaStream_t stream);
// d_order[rank] = base + i for the `shots` entries of d_cost, descending cost, ties by index (stable): the order
// of a wave's second launch, formed in-stream so that no host round trip separates the two launches
cudaError_t lap_order_by_cost(const float* d_cost, int* d_order, int shots, int base, cudaStream_t stream);

}  // namespace pqb200
