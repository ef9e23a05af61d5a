// laplace_sampler.cuh -- the Clifford-Clifford "B" sampler's photon loop, entirely on the GPU:
// ONE CTA per shot walks through all photon steps of fill_r_sample
// (reference: sampling/simulation_strategies/source/CGeneralizedCliffordsBSimulationStrategy.cpp:282-316,
//  samplingHelperFunctions.cpp:78-342 compute_pmf, :350-374 sample_from_pmf).
//
// What one photon step computes.  With `sample` (K = k-1 photons already placed in output modes)
// and `cur` (k photons inserted so far) the reference evaluates, for every occupied input mode i,
//     addend_i = perm( U[rows = sample, cols = cur - e_i] )          (k separate repeated permanents)
// and then pmf[m'] ~ | sum_i cur_i * addend_i * U[m'][i] |^2.
// All k sub-permanents share the row side.  Writing Glynn's formula over the ROW multiplicities,
//     addend_i = 2^-(K-1) sum_{delta} w(delta) prod_j s_j(delta)^{c_j - [j == i]},
// every term of the Gray sweep is the same vector of column sums s_j with ONE factor left out
// (Laplace expansion / "Lemma 1" of Clifford & Clifford; piquasso-module/src/permanent_laplace.cpp:43-233
// is the reference's CPU form).  So one sweep with prefix/suffix products yields all k addends:
// ~14 N FP64 instructions per term instead of k * 6 N.
//
// GPU layout.
//   * One thread-block CLUSTER (1, 2, 4 or 8 CTAs of 128 threads) per shot, all photon steps in one
//     launch: insertion order and the uniform draws of the shot are pre-resolved on the host from the
//     reference's std::mt19937 stream (they do not depend on the pmf values), so there is no host round
//     trip between steps.  The CTAs of a cluster split the chunks of every Gray sweep and combine their
//     partial addends through distributed shared memory in a fixed order.  Every CTA of a cluster repeats
//     the step's preparation, so the host only uses clusters to fill otherwise idle SMs (a 125-shot shard
//     of an 8-GPU run would occupy a quarter of them).
//   * A shot's cost is set by the collisions in its sample, which nobody knows in advance, and a wave of
//     1000 shots on 444 resident CTAs is only 2.25 shots per CTA: handed out blindly, the last CTAs run
//     alone for a third of the launch.  So a wave is TWO launches: all photon steps but the last three,
//     then -- with the shots re-ordered by the size of the Gray space they have reached, longest first --
//     the last three steps, which hold 7/8 of the arithmetic (each step costs twice the one before).
//   * Per step the CTA builds the job in shared memory (rows / radices from `sample`, physically
//     repeated columns from `cur`, gather of the sub-matrix from U = adaptInterferometer), the
//     low-digit step table and the two base vectors -- same low/high digit split of the reflected
//     mixed-radix Gray code as glynn_repeated.cuh.
//   * A Gray term is owned by LP in {1,2,4} adjacent lanes; each lane keeps NC column sums, NC suffix
//     products and NC accumulators in registers, the lanes exchange their partial products with one
//     or two shuffles.  (NC, LP) is picked per step from the photon count.
//   * Accumulators are reduced by shuffles, then across the 4 warps in a fixed order; threads then
//     form the m amplitudes; thread 0 normalises and draws exactly as sample_from_pmf does.
// The same kernel serves pqb200_pmf_layer (mode 1: `sample`/`cur` given, one step, pmf written out).
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "glynn_plain.cuh"
#include "glynn_repeated.cuh"

namespace pqb200 {

constexpr int kLapThreads = 128;
constexpr int kLapMaxPhotons = 40;  // columns of a step (photons inserted so far)
constexpr int kLapMaxLen = 512;     // low-digit steps per chunk
constexpr int kLapBinomDim = 48;
constexpr int kLapMaxCluster = 8;   // CTAs that share one shot (portable cluster size limit)

static __constant__ double c_lap_binom[kLapBinomDim][kLapBinomDim];

// PQB_LAP_PROFILE (tuning builds only): thread 0 of every CTA accumulates the cycles between phase
// boundaries into LapArgs::prof[phase] (tools/sampler_phases.py prints the shares)
#ifdef PQB_LAP_PROFILE
#define LAP_TICK(ph)                                                                   \
    do {                                                                               \
        if (threadIdx.x == 0 && a.prof) {                                              \
            const long long now__ = clock64();                                         \
            atomicAdd(a.prof + (ph), (unsigned long long)(now__ - lap_t0__));          \
            lap_t0__ = now__;                                                          \
        }                                                                              \
    } while (0)
#define LAP_TICK_DECL long long lap_t0__ = clock64()
#else
#define LAP_TICK(ph) do {} while (0)
#define LAP_TICK_DECL do {} while (0)
#endif

struct LapArgs {
    const double2* U;  // m x m interferometer, row stride `ustride` complex elements
    int ustride;
    int m;
    int shots;
    int nmax;                        // row length of ins / rnd
    const int* order;                // [shots] shot handled by cluster i (descending cost), NULL = identity
    int chunks_per_owner;            // chunks per term owner the low/high digit split aims at
    unsigned long long* prof;        // PQB_LAP_PROFILE builds: [16] cycle counters per phase (else unused)
    // mode 0 may be run as several launches over consecutive ranges of photon steps (the light early
    // steps one CTA per shot, the heavy last steps split over a cluster): this launch does steps
    // [step_begin, step_end) of every shot; `state` ([shots][2][m] bytes: sample, cur) carries a shot
    // from one launch to the next
    int step_begin, step_end;
    unsigned char* state;
    float* cost;                     // [shots] Gray terms of the last step a handed-over shot has done: the
                                     // next launch is ordered by it, longest first (lap_rank_kernel)
    const int* nphot;                // [shots] photon steps of the shot                    (mode 0)
    const unsigned char* ins;        // [shots][nmax] input mode of the k-th inserted photon (mode 0)
    const double* rnd;               // [shots][nmax] uniform draw of step k                 (mode 0)
    long long* samples_out;          // [shots][m]                                           (mode 0)
    const unsigned char* sample_in;  // [shots][m]                                           (mode 1)
    const unsigned char* cur_in;     // [shots][m]                                           (mode 1)
    double* pmf_out;                 // [shots][m]                                           (mode 1)
    unsigned long long* gray_terms;  // statistics: Gray terms swept by the launch (atomicAdd per step)
    int mode;
};

struct LapShared {
    double2 A[kLapMaxPhotons][kLapMaxPhotons];  // sub-matrix of the step: row 0 = anchor, columns repeated
    double2 base[2][kLapMaxPhotons];
    RepStep table[kLapMaxLen];
    double2 red[kLapThreads / 32][kLapMaxPhotons];
    double2 cred[kLapMaxCluster][kLapMaxPhotons];  // rank 0: the partial addends of every CTA of the cluster
    double2 addend[kLapMaxPhotons];
    double p[256];
    double tot;
    unsigned long long highW[kLapMaxPhotons];
    unsigned long long nchunks;
    unsigned int lowW[kLapMaxPhotons];
    unsigned int len;
    int lowPar[2];
    int R, N, q, K, chosen;
    unsigned char sample[256], cur[256];
    double rnd[kLapMaxPhotons];        // the shot's uniform draws and insertion order, fetched once (mode 0):
    unsigned char ins[kLapMaxPhotons]; // a global load per photon step would sit on the critical path
    unsigned char gidx[kLapMaxPhotons], radix[kLapMaxPhotons], pidx[kLapMaxPhotons], first[kLapMaxPhotons];
};

__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// lanes per Gray term for a step with N columns
__host__ __device__ __forceinline__ int lap_lanes(int N) { return N <= 10 ? 1 : N <= 20 ? 2 : 4; }

// Instantiation of a step with N columns: (NC columns per lane, LP lanes per term), NC * LP >= N (the
// padding columns hold the constant 1).  The light early steps share ONE instantiation per lane count
// -- (10,1) for N <= 10, (8,2) for N <= 16: with three CTAs per SM in different photon steps, a separate
// instantiation per column count thrashed the instruction cache (ncu: stall "no_instruction" 1.8 per
// issue over the launch of the light steps, profiles/ncu_sampler_c5_r2.md); the extra columns cost
// nothing there (<= 2^15 Gray terms).  From N = 17 on NC is the exact ceil(N / LP).
__host__ __device__ __forceinline__ int lap_bucket(int N) {
    if (N <= 10) return 5;    // (10,1)
    if (N <= 16) return 13;   // (8,2)
    if (N <= 20) return 5 + (N + 1) / 2;    // (9,2), (10,2)
    return 15 + (N + 3) / 4;  // (6,4) .. (10,4)
}

// One Gray sweep of the current step: fills sh.addend[0..N) (unscaled by 2^-(K-1)).
template <int NC, int LP>
__device__ __noinline__ void lap_sweep(LapShared& sh, const unsigned int crank, const unsigned int csize) {
    constexpr int G = kLapThreads / LP;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = tid / LP, c0 = (tid % LP) * NC;
    const int D = sh.R - 1, q = sh.q;
    const unsigned int len = sh.len;
    const unsigned long long nchunks = sh.nchunks;
    const bool wide = nchunks > 0xffffffffull;
    double ar[NC], ai[NC];
#pragma unroll
    for (int j = 0; j < NC; j++) ar[j] = ai[j] = 0.0;
    const unsigned long long per_it = (unsigned long long)G * csize;  // chunks the cluster takes per iteration
    const unsigned long long iters = (nchunks + per_it - 1) / per_it;
    for (unsigned long long it = 0; it < iters; it++) {
        const unsigned long long c = it * per_it + (unsigned long long)crank * G + (unsigned long long)g;
        const bool live = c < nchunks;
        const unsigned long long cc = live ? c : 0ull;
        // high digits (most significant first): reflected Gray values, parity, binomial weight and
        // the high rows' contribution to the column sums (n_aryGrayCodeCounter.cpp:114-140)
        int par = 0, ksum = 0;
        double whigh = 1.0;
        double sr[NC], si[NC];
#pragma unroll
        for (int j = 0; j < NC; j++) sr[j] = si[j] = 0.0;
        for (int d = D - 1; d >= q; d--) {
            const int radix = sh.radix[d];
            const int cnt = wide ? (int)((cc / sh.highW[d]) % (unsigned long long)radix)
                                 : (int)(((unsigned int)cc / (unsigned int)sh.highW[d]) % (unsigned int)radix);
            const int k = par ? radix - 1 - cnt : cnt;
            par ^= k & 1;
            ksum += k;
            whigh *= c_lap_binom[radix - 1][k];
            const double coef = (double)(radix - 1 - 2 * k);
#pragma unroll
            for (int j = 0; j < NC; j++) {
                const double2 a = sh.A[d + 1][c0 + j];
                sr[j] = fma(coef, a.x, sr[j]);
                si[j] = fma(coef, a.y, si[j]);
            }
        }
        const int P = par;
#pragma unroll
        for (int j = 0; j < NC; j++) {
            const double2 b = sh.base[P][c0 + j];
            sr[j] += b.x;
            si[j] += b.y;
        }
        const double f = live ? ((((ksum + sh.lowPar[P]) & 1) ? -whigh : whigh)) : 0.0;
        for (unsigned int t = 0; t < len; t++) {
            const RepStep e = sh.table[t];
            if (t) {
                const double sg = ((e.flags >> P) & 1u) ? -2.0 : 2.0;
#pragma unroll
                for (int j = 0; j < NC; j++) {
                    const double2 a = sh.A[e.row][c0 + j];
                    sr[j] = fma(sg, a.x, sr[j]);
                    si[j] = fma(sg, a.y, si[j]);
                }
            }
            // suffix products fr[j] = prod_{l >= j} s_l of this lane's columns
            double fr[NC], fi[NC];
            fr[NC - 1] = sr[NC - 1];
            fi[NC - 1] = si[NC - 1];
#pragma unroll
            for (int j = NC - 2; j >= 0; j--) {
                fr[j] = fr[j + 1];
                fi[j] = fi[j + 1];
                cmul(fr[j], fi[j], sr[j], si[j]);
            }
            // running prefix, started at weight * (product of the OTHER lanes' columns)
            const double w = e.w * f;
            double pr, pi;
            if (LP == 1) {
                pr = w;
                pi = 0.0;
            } else if (LP == 2) {
                pr = w * shfl_xor_d(fr[0], 1);
                pi = w * shfl_xor_d(fi[0], 1);
            } else {
                double xr = shfl_xor_d(fr[0], 1), xi = shfl_xor_d(fi[0], 1);  // neighbour's product
                double tr = fr[0], ti = fi[0];
                cmul(tr, ti, xr, xi);                                        // product of my pair
                const double yr = shfl_xor_d(tr, 2), yi = shfl_xor_d(ti, 2);  // the other pair's product
                cmul(xr, xi, yr, yi);
                pr = w * xr;
                pi = w * xi;
            }
#pragma unroll
            for (int j = 0; j < NC - 1; j++) {
                cmul_acc<false>(ar[j], ai[j], pr, pi, fr[j + 1], fi[j + 1]);
                cmul(pr, pi, sr[j], si[j]);
            }
            ar[NC - 1] += pr;
            ai[NC - 1] += pi;
        }
    }
    // ---- reduce over the Gray terms: lanes with the same column range, then the warps (fixed order)
#pragma unroll
    for (int off = LP; off < 32; off <<= 1) {
#pragma unroll
        for (int j = 0; j < NC; j++) {
            ar[j] += shfl_xor_d(ar[j], off);
            ai[j] += shfl_xor_d(ai[j], off);
        }
    }
    if (lane < LP) {
#pragma unroll
        for (int j = 0; j < NC; j++)
            if (c0 + j < kLapMaxPhotons) sh.red[warp][c0 + j] = make_double2(ar[j], ai[j]);
    }
    __syncthreads();
    if (tid < sh.N) {
        double re = 0.0, im = 0.0;
#pragma unroll
        for (int w = 0; w < kLapThreads / 32; w++) {
            re += sh.red[w][tid].x;
            im += sh.red[w][tid].y;
        }
        sh.addend[tid] = make_double2(re, im);
    }
    __syncthreads();
    if (csize > 1) {
        // cluster reduction through distributed shared memory: every CTA stores its partial addends
        // into rank 0's `cred`, rank 0 folds them in rank order (deterministic)
        namespace cg = cooperative_groups;
        cg::cluster_group cluster = cg::this_cluster();
        LapShared* sh0 = cluster.map_shared_rank(&sh, 0);
        if (tid < sh.N) sh0->cred[crank][tid] = sh.addend[tid];
        cluster.sync();
        if (crank == 0 && tid < sh.N) {
            double re = 0.0, im = 0.0;
            for (unsigned int r = 0; r < csize; r++) {
                re += sh.cred[r][tid].x;
                im += sh.cred[r][tid].y;
            }
            sh.addend[tid] = make_double2(re, im);
        }
        __syncthreads();
    }
}

// Build the step's job in shared memory from sh.sample / sh.cur (thread 0), then gather and tabulate
// (all threads).  Returns with everything lap_sweep needs in place.
#ifdef PQB_LAP_PROFILE
#define LAP_T0_PARAM , long long& lap_t0__
#define LAP_T0_ARG , lap_t0__
#else
#define LAP_T0_PARAM
#define LAP_T0_ARG
#endif
__device__ __forceinline__ void lap_prepare(LapShared& sh, const LapArgs& a, const unsigned int csize LAP_T0_PARAM) {
    const int tid = threadIdx.x, m = a.m;
    if (tid < 32) {
        // Warp 0 builds the job cooperatively: lane L owns the modes [8L, 8L+8) (m <= 255), positions in
        // the row / column lists come from warp prefix sums, so the lists keep the reference's mode order.
        const unsigned int FULL = 0xffffffffu;
        const int lo = 8 * tid;
        // rows: occupied output modes; the anchor takes one copy of the FIRST minimum-multiplicity row
        // (BBFGPermanentCalculatorRepeated.hpp:78-94): minimise (multiplicity, index)
        int K = 0;
        unsigned int key = 0xffffffffu;
        unsigned char smp[8], cu[8];
#pragma unroll
        for (int b = 0; b < 8; b++) {
            const int i = lo + b;
            smp[b] = i < m ? sh.sample[i] : 0;
            cu[b] = i < m ? sh.cur[i] : 0;
            K += smp[b];
            if (smp[b] > 0) {
                const unsigned int k2 = ((unsigned int)smp[b] << 8) | (unsigned int)i;
                key = k2 < key ? k2 : key;
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            K += __shfl_xor_sync(FULL, K, off);
            const unsigned int o = __shfl_xor_sync(FULL, key, off);
            key = o < key ? o : key;
        }
        const int amin = key == 0xffffffffu ? -1 : (int)(key & 0xffu);
        int nrow = 0, ncol = 0;
#pragma unroll
        for (int b = 0; b < 8; b++) {
            nrow += ((int)smp[b] - (lo + b == amin ? 1 : 0)) > 0 ? 1 : 0;
            ncol += cu[b];
        }
        int prow = nrow, pcol = ncol;  // inclusive prefix sums over the lanes
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int a1 = __shfl_up_sync(FULL, prow, off), a2 = __shfl_up_sync(FULL, pcol, off);
            if (tid >= off) {
                prow += a1;
                pcol += a2;
            }
        }
        const int R = amin >= 0 ? 1 + __shfl_sync(FULL, prow, 31) : 0;
        int Ntot = __shfl_sync(FULL, pcol, 31);
        Ntot = Ntot > kLapMaxPhotons ? kLapMaxPhotons : Ntot;
        if (amin >= 0) {
            if (tid == 0) sh.gidx[0] = (unsigned char)amin;
            int pos = 1 + prow - nrow;
#pragma unroll
            for (int b = 0; b < 8; b++) {
                const int ml = (int)smp[b] - (lo + b == amin ? 1 : 0);
                if (ml > 0) {
                    sh.radix[pos - 1] = (unsigned char)(ml + 1);
                    sh.gidx[pos] = (unsigned char)(lo + b);
                    pos++;
                }
            }
        }
        {   // columns: occupied input modes, physically repeated
            int pos = pcol - ncol;
#pragma unroll
            for (int b = 0; b < 8; b++)
                for (int rep = 0; rep < (int)cu[b] && pos < kLapMaxPhotons; rep++) {
                    sh.pidx[pos] = (unsigned char)(lo + b);
                    sh.first[pos] = rep == 0;
                    pos++;
                }
        }
        __syncwarp();
        // total Gray terms = prod radix (at most 40 digits: two per lane)
        double total = 1.0;
        if (tid < R - 1) total *= (double)sh.radix[tid];
        if (tid + 32 < R - 1) total *= (double)sh.radix[tid + 32];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) total *= __shfl_xor_sync(FULL, total, off);
        if (tid == 0) {
            sh.R = R;
            sh.N = Ntot;
            sh.K = K;
            // low / high digit split: chunks of at most kLapMaxLen steps, about chunks_per_owner chunks per
            // term owner of the cluster (the last round of chunks is partly idle)
            const int G = kLapThreads / lap_lanes(Ntot) * (int)csize;
            double want = total / ((double)a.chunks_per_owner * G);
            want = want < 1.0 ? 1.0 : want > (double)kLapMaxLen ? (double)kLapMaxLen : want;
            unsigned int len = 1;
            int q = 0;
            while (q < R - 1 && (double)len * sh.radix[q] <= want) {
                sh.lowW[q] = len;
                len *= sh.radix[q];
                q++;
            }
            unsigned long long nch = 1;
            for (int d = q; d < R - 1; d++) {
                sh.highW[d] = nch;
                nch *= sh.radix[d];
            }
            sh.q = q;
            sh.len = len;
            sh.nchunks = nch;
        }
    }
    __syncthreads();
    LAP_TICK(0);  // job construction (warp 0)
    const int R = sh.R, N = sh.N, q = sh.q;
    const unsigned int len = sh.len;
    // gather (adaptInterferometer, BatchedPermanentCalculator.cpp:412-451): R x N loads from U; the
    // padding columns (the sweep works on whole groups of NC columns per lane) hold the constant 1 in
    // the anchor row so that their column sum is 1 for every term
    for (int e = tid; e < R * N; e += kLapThreads) {
        const int i = e / N, j = e - i * N;
        sh.A[i][j] = a.U[(size_t)sh.gidx[i] * a.ustride + sh.pidx[j]];
    }
    for (int e = tid; e < R * (kLapMaxPhotons - N); e += kLapThreads) {
        const int i = e / (kLapMaxPhotons - N), j = N + e - i * (kLapMaxPhotons - N);
        sh.A[i][j] = make_double2(i == 0 ? 1.0 : 0.0, 0.0);
    }
#ifdef PQB_LAP_PROFILE
    __syncthreads();
#endif
    LAP_TICK(1);  // gather
    // low-digit step table: changed digit, direction for both high parities, signed binomial weight.
    // The counter digits of t come from ONE chain of divisions (least significant digit first); the
    // changed digit is the lowest non-zero one, and the reflected Gray values for both parities of the
    // high digits follow from a single pass from the most significant digit down.
    for (unsigned int t = tid; t < len; t += kLapThreads) {
        unsigned long long pk = 0ull;  // counter digits, 6 bits each (q <= log2(kLapMaxLen) = 9 digits < 41)
        unsigned int rem = t;
        int ch = 0;
        bool found = false;
#pragma unroll 1
        for (int d = 0; d < q; d++) {
            const unsigned int rd = sh.radix[d];
            const unsigned int qq = rem / rd;
            const unsigned int cd = rem - qq * rd;
            pk |= (unsigned long long)cd << (6 * d);
            rem = qq;
            if (!found && cd != 0) {
                ch = d;
                found = true;
            }
        }
        unsigned int flags = 0;
        double w = 1.0;
        int par0 = 0, par1 = 1;
#pragma unroll 1
        for (int d = q - 1; d >= 0; d--) {
            const int r = sh.radix[d] - 1;
            const int c = (int)((pk >> (6 * d)) & 63ull);
            const int g0 = par0 ? r - c : c, g1 = par1 ? r - c : c;
            if (d == ch) flags = (par0 == 0 ? 1u : 0u) | (par1 == 0 ? 2u : 0u);
            par0 ^= g0 & 1;
            par1 ^= g1 & 1;
            w *= c_lap_binom[r][g0];
        }
        RepStep e;
        e.w = (t & 1u) ? -w : w;
        e.row = ch + 1;
        e.flags = flags;
        sh.table[t] = e;
    }
    __syncthreads();
    LAP_TICK(2);  // step table
    // base vectors: anchor + low rows at their initial Gray value, for both high parities
    for (int e = tid; e < 2 * kLapMaxPhotons; e += kLapThreads) {
        const int P = e / kLapMaxPhotons, j = e % kLapMaxPhotons;
        double re = sh.A[0][j].x, im = sh.A[0][j].y;
        int par = P, ks = 0;
        for (int d = q - 1; d >= 0; d--) {
            const int r = sh.radix[d] - 1;
            const int gk = par ? r : 0;
            par ^= gk & 1;
            ks += gk;
            re += (double)(r - 2 * gk) * sh.A[d + 1][j].x;
            im += (double)(r - 2 * gk) * sh.A[d + 1][j].y;
        }
        sh.base[P][j] = make_double2(re, im);
        if (j == 0) sh.lowPar[P] = ks & 1;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kLapThreads, 3) laplace_sampler_kernel(const LapArgs a) {
    namespace cg = cooperative_groups;
    __shared__ LapShared sh;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned int csize = cluster.num_blocks(), crank = cluster.block_rank();
    // every CTA of the cluster must be running before any of them touches another one's shared memory
    // (the first photon step has no sweep, hence no barrier of its own, before rank 0 publishes the draw)
    if (csize > 1) cluster.sync();
    const int tid = threadIdx.x, m = a.m;
    const int slot = (int)(blockIdx.x / csize);
    const int s = a.order ? a.order[slot] : slot;
    int steps;
    LAP_TICK_DECL;
    int k_first = 0;
    if (a.mode == 0) {
        const int np = a.nphot[s];
        k_first = a.step_begin < np ? a.step_begin : np;
        steps = a.step_end < np ? a.step_end : np;
        if (k_first >= steps && a.step_begin > 0) return;  // the shot was finished by an earlier launch (uniform per cluster)
        if (tid < steps) {
            sh.rnd[tid] = a.rnd[(size_t)s * a.nmax + tid];
            sh.ins[tid] = a.ins[(size_t)s * a.nmax + tid];
        }
        if (k_first == 0) {
            for (int i = tid; i < m; i += kLapThreads) sh.sample[i] = sh.cur[i] = 0;
        } else {
            const unsigned char* st = a.state + (size_t)s * 2 * m;
            for (int i = tid; i < m; i += kLapThreads) {
                sh.sample[i] = st[i];
                sh.cur[i] = st[m + i];
            }
        }
    } else {
        for (int i = tid; i < m; i += kLapThreads) {
            sh.sample[i] = a.sample_in[(size_t)s * m + i];
            sh.cur[i] = a.cur_in[(size_t)s * m + i];
        }
        steps = 1;
    }
    __syncthreads();
    // every CTA of the cluster keeps its own copy of the shot's state (sample, cur) and prepares each
    // step redundantly; only the Gray sweep is split
    for (int k = k_first; k < steps; k++) {
        if (a.mode == 0) {
            // update_input_by_single_photon (samplingHelperFunctions.cpp:377-406), order pre-resolved
            if (tid == 0) sh.cur[sh.ins[k]]++;
            __syncthreads();
        }
        lap_prepare(sh, a, csize LAP_T0_ARG);
        LAP_TICK(3);  // base vectors
        const int N = sh.N;
        if (tid == 0 && crank == 0 && sh.R > 0) atomicAdd(a.gray_terms, (unsigned long long)sh.len * sh.nchunks);
        if (sh.R == 0) {
            // no photon placed yet: every sub-permanent is the empty permanent 1 (hpp:126-128)
            if (tid < N) sh.addend[tid] = make_double2(1.0, 0.0);
            __syncthreads();
        } else {
            switch (lap_bucket(N)) {
                case 5: lap_sweep<10, 1>(sh, crank, csize); break;   // N <= 10
                case 13: lap_sweep<8, 2>(sh, crank, csize); break;   // N = 11..16
                case 14: lap_sweep<9, 2>(sh, crank, csize); break;
                case 15: lap_sweep<10, 2>(sh, crank, csize); break;
                // four lanes per term: N = 21..40 -> NC = 6..10
                case 21: lap_sweep<6, 4>(sh, crank, csize); break;
                case 22: lap_sweep<7, 4>(sh, crank, csize); break;
                case 23: lap_sweep<8, 4>(sh, crank, csize); break;
                case 24: lap_sweep<9, 4>(sh, crank, csize); break;
                default: lap_sweep<10, 4>(sh, crank, csize); break;
            }
        }
        LAP_TICK(4 + (sh.K < 8 ? 0 : sh.K < 12 ? 1 : sh.K < 16 ? 2 : 3));  // sweep, by photons placed
        // ---- amplitudes and probabilities of the m output modes (samplingHelperFunctions.cpp:308-339):
        //      rank 0 of the cluster holds the complete addends
        if (crank == 0) {
            // All threads: output mode mm = tid % m, the occupied input modes dealt round-robin over the
            // kLapThreads / m thread groups (fixed trip count, predicated: the loads of U are independent and
            // issue back to back); partial amplitudes are combined through shared memory.
            const double sc = sh.K > 0 ? ldexp(1.0, -(sh.K - 1)) : 1.0;
            const int groups = m > 64 ? 1 : (kLapThreads / m > 4 ? 4 : kLapThreads / m);
            double2* const amp = &sh.cred[0][0];  // [4][64] partial amplitudes; `cred` is free by now
            static_assert(kLapMaxCluster * kLapMaxPhotons >= 4 * 64, "amplitude scratch aliases cred");
            if (groups > 1) {
                const int mm = tid % m, grp = tid / m;
                double re = 0.0, im = 0.0;
                if (grp < groups) {
#pragma unroll
                    for (int it = 0; it < (kLapMaxPhotons + 1) / 2; it++) {  // groups >= 2: at most 20 columns per group
                        const int j = grp + it * groups;
                        if (j < N && sh.first[j]) {
                            const int col = sh.pidx[j];
                            const double cw = (double)sh.cur[col] * sc;
                            const double2 ad = sh.addend[j];
                            const double2 u = a.U[(size_t)mm * a.ustride + col];
                            const double wr = cw * ad.x, wi = cw * ad.y;
                            re += wr * u.x - wi * u.y;
                            im += wr * u.y + wi * u.x;
                        }
                    }
                    amp[grp * 64 + mm] = make_double2(re, im);
                }
                __syncthreads();
                if (tid < m) {
                    re = 0.0;
                    im = 0.0;
                    for (int g2 = 0; g2 < groups; g2++) {
                        re += amp[g2 * 64 + tid].x;
                        im += amp[g2 * 64 + tid].y;
                    }
                    sh.p[tid] = re * re + im * im;
                }
            } else {
                for (int mm = tid; mm < m; mm += kLapThreads) {
                    double re = 0.0, im = 0.0;
                    for (int j = 0; j < N; j++) {
                        if (!sh.first[j]) continue;
                        const int col = sh.pidx[j];
                        const double cw = (double)sh.cur[col] * sc;
                        const double wr = cw * sh.addend[j].x, wi = cw * sh.addend[j].y;
                        const double2 u = a.U[(size_t)mm * a.ustride + col];
                        re += wr * u.x - wi * u.y;
                        im += wr * u.y + wi * u.x;
                    }
                    sh.p[mm] = re * re + im * im;
                }
            }
            __syncthreads();
            // Warp 0: probability_sum, the normalised layer and the draw (sample_from_pmf,
            // samplingHelperFunctions.cpp:350-374: the first index whose cumulative normalised probability
            // reaches the draw, 0 if none does).  Lane L owns the modes [8L, 8L+8); the cumulative sums come
            // from a warp prefix scan, i.e. in another association order than the reference's running sum
            // (differences of an ulp, far below the kernel's own rounding).
            if (tid < 32) {
                const unsigned int FULL = 0xffffffffu;
                double pv[8], part = 0.0;
#pragma unroll
                for (int b = 0; b < 8; b++) {
                    const int i = 8 * tid + b;
                    pv[b] = i < m ? sh.p[i] : 0.0;
                    part += pv[b];
                }
                double tot = part;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) tot += __shfl_xor_sync(FULL, tot, off);
                if (a.mode == 0) {
                    double run = 0.0;
#pragma unroll
                    for (int b = 0; b < 8; b++) {
                        pv[b] = pv[b] / tot;  // the normalised layer
                        run += pv[b];
                    }
                    double incl = run;  // inclusive prefix of the lanes' sums
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        const double o = __shfl_up_sync(FULL, incl, off);
                        if (tid >= off) incl += o;
                    }
                    const double r = sh.rnd[k];
                    double acc = incl - run;
                    int hit = -1;
#pragma unroll
                    for (int b = 0; b < 8; b++) {
                        acc += pv[b];
                        if (hit < 0 && 8 * tid + b < m && acc >= r) hit = 8 * tid + b;
                    }
                    const unsigned int any = __ballot_sync(FULL, hit >= 0);
                    int chosen = 0;
                    if (any) chosen = __shfl_sync(FULL, hit, __ffs(any) - 1);
                    if (tid == 0)
                        for (unsigned int r2 = 0; r2 < csize; r2++) cluster.map_shared_rank(&sh, r2)->chosen = chosen;
                } else if (tid == 0) {
                    sh.tot = tot;
                }
            }
        }
        if (csize > 1) cluster.sync(); else __syncthreads();
        if (a.mode == 0) {
            if (tid == 0) sh.sample[sh.chosen]++;
            __syncthreads();
        }
        LAP_TICK(8);  // amplitudes, pmf, draw
    }
    if (crank == 0) {
        if (a.mode == 0) {
            if (steps >= a.nphot[s]) {
                for (int i = tid; i < m; i += kLapThreads) a.samples_out[(size_t)s * m + i] = (long long)sh.sample[i];
            } else {  // hand the shot over to the next launch
                if (tid == 0) a.cost[s] = sh.R > 0 ? (float)sh.len * (float)sh.nchunks : 1.0f;
                unsigned char* st = a.state + (size_t)s * 2 * m;
                for (int i = tid; i < m; i += kLapThreads) {
                    st[i] = sh.sample[i];
                    st[m + i] = sh.cur[i];
                }
            }
        } else {
            const double tot = sh.tot;
            for (int i = tid; i < m; i += kLapThreads) a.pmf_out[(size_t)s * m + i] = sh.p[i] / tot;
        }
    }
}

// Order of the second launch of a wave, formed on the device so that the host can queue both launches of a
// wave back to back: order[rank] = base + i with the shots ranked by descending cost, ties by ascending index
// (what a stable sort gives).  S is a few thousand at most: every thread counts the shots ahead of its own
// against shared-memory tiles of the cost vector.
constexpr int kLapRankThreads = 256;
__global__ void __launch_bounds__(kLapRankThreads) lap_rank_kernel(const float* __restrict__ cost, int* __restrict__ order,
                                                                   const int S, const int base) {
    __shared__ float tile[kLapRankThreads];
    const int i = (int)(blockIdx.x * kLapRankThreads + threadIdx.x);
    const float mine = i < S ? cost[i] : 0.0f;
    int rank = 0;
    for (int j0 = 0; j0 < S; j0 += kLapRankThreads) {
        const int j = j0 + (int)threadIdx.x;
        tile[threadIdx.x] = j < S ? cost[j] : -1.0f;  // costs are >= 0: the padding ranks behind everything
        __syncthreads();
        const int n = S - j0 < kLapRankThreads ? S - j0 : kLapRankThreads;
        for (int t = 0; t < n; t++) {
            const float c = tile[t];
            rank += (c > mine || (c == mine && j0 + t < i)) ? 1 : 0;
        }
        __syncthreads();
    }
    if (i < S) order[rank] = base + i;
}

}  // namespace pqb200
