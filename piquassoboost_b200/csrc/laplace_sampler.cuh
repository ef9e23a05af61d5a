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
//     partial addends through distributed shared memory in a fixed order; the host picks the cluster
//     size so that shots x cluster covers the GPU several times (a 125-shot shard of an 8-GPU run would
//     otherwise occupy a quarter of the SMs) and hands out the shots in descending photon count.
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

struct LapArgs {
    const double2* U;  // m x m interferometer, row stride `ustride` complex elements
    int ustride;
    int m;
    int shots;
    int nmax;                        // row length of ins / rnd
    const int* order;                // [shots] shot handled by cluster i (descending cost), NULL = identity
    int chunks_per_owner;            // chunks per term owner the low/high digit split aims at
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
    double2 wadd[kLapMaxPhotons];   // weighted addends of the distinct occupied input modes ...
    double p[256];
    double tot;
    int nd;                         // ... and how many there are
    unsigned char wcol[kLapMaxPhotons];
    unsigned long long highW[kLapMaxPhotons];
    unsigned long long nchunks;
    unsigned int lowW[kLapMaxPhotons];
    unsigned int len;
    int lowPar[2];
    int R, N, q, K, chosen;
    unsigned char sample[256], cur[256];
    unsigned char gidx[kLapMaxPhotons], radix[kLapMaxPhotons], pidx[kLapMaxPhotons], first[kLapMaxPhotons];
};

__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// lanes per Gray term for a step with N columns
__host__ __device__ __forceinline__ int lap_lanes(int N) { return N <= 10 ? 1 : N <= 20 ? 2 : 4; }

// instantiation id of a step with N columns: LP = 1 -> NC = 2,4,..,10 (ids 1..5); LP = 2 -> NC = 6..10
// (ids 11..15); LP = 4 -> NC = 6..10 (ids 21..25)
__host__ __device__ __forceinline__ int lap_bucket(int N) {
    const int lp = lap_lanes(N);
    return lp == 1 ? (N + 1) / 2 : lp == 2 ? 5 + (N + 1) / 2 : 15 + (N + 3) / 4;
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
__device__ __forceinline__ void lap_prepare(LapShared& sh, const LapArgs& a, const unsigned int csize) {
    const int tid = threadIdx.x, m = a.m;
    if (tid == 0) {
        // rows: occupied output modes; the anchor takes one copy of the first minimum-multiplicity
        // row (BBFGPermanentCalculatorRepeated.hpp:78-94)
        int K = 0, amin = -1, mn = 1 << 30;
        for (int i = 0; i < m; i++) {
            const int v = sh.sample[i];
            if (v > 0) {
                K += v;
                if (v < mn) { mn = v; amin = i; }
            }
        }
        int R = 0;
        if (amin >= 0) {
            sh.gidx[R++] = (unsigned char)amin;
            for (int i = 0; i < m; i++) {
                const int ml = (int)sh.sample[i] - (i == amin ? 1 : 0);
                if (ml > 0) {
                    sh.radix[R - 1] = (unsigned char)(ml + 1);
                    sh.gidx[R++] = (unsigned char)i;
                }
            }
        }
        // columns: occupied input modes, physically repeated
        int N = 0;
        for (int i = 0; i < m; i++)
            for (int rep = 0; rep < (int)sh.cur[i] && N < kLapMaxPhotons; rep++) {
                sh.pidx[N] = (unsigned char)i;
                sh.first[N] = rep == 0;
                N++;
            }
        sh.R = R;
        sh.N = N;
        sh.K = K;
        // low / high digit split: chunks of at most kLapMaxLen steps, >= 8 chunks per term owner of the
        // cluster (the last round of chunks is partly idle: at most 1/8 of the sweep)
        double total = 1.0;
        for (int d = 0; d < R - 1; d++) total *= (double)sh.radix[d];
        const int G = kLapThreads / lap_lanes(N) * (int)csize;
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
    __syncthreads();
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
    const int tid = threadIdx.x, m = a.m;
    const int slot = (int)(blockIdx.x / csize);
    const int s = a.order ? a.order[slot] : slot;
    int steps;
    if (a.mode == 0) {
        for (int i = tid; i < m; i += kLapThreads) sh.sample[i] = sh.cur[i] = 0;
        steps = a.nphot[s];
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
    for (int k = 0; k < steps; k++) {
        if (a.mode == 0) {
            // update_input_by_single_photon (samplingHelperFunctions.cpp:377-406), order pre-resolved
            if (tid == 0) sh.cur[a.ins[(size_t)s * a.nmax + k]]++;
            __syncthreads();
        }
        lap_prepare(sh, a, csize);
        const int N = sh.N;
        if (tid == 0 && crank == 0 && sh.R > 0) atomicAdd(a.gray_terms, (unsigned long long)sh.len * sh.nchunks);
        if (sh.R == 0) {
            // no photon placed yet: every sub-permanent is the empty permanent 1 (hpp:126-128)
            if (tid < N) sh.addend[tid] = make_double2(1.0, 0.0);
            __syncthreads();
        } else {
            switch (lap_bucket(N)) {
                // one lane per term: NC = 2,4,..,10
                case 1: lap_sweep<2, 1>(sh, crank, csize); break;
                case 2: lap_sweep<4, 1>(sh, crank, csize); break;
                case 3: lap_sweep<6, 1>(sh, crank, csize); break;
                case 4: lap_sweep<8, 1>(sh, crank, csize); break;
                case 5: lap_sweep<10, 1>(sh, crank, csize); break;
                // two lanes per term: N = 11..20 -> NC = 6..10
                case 11: lap_sweep<6, 2>(sh, crank, csize); break;
                case 12: lap_sweep<7, 2>(sh, crank, csize); break;
                case 13: lap_sweep<8, 2>(sh, crank, csize); break;
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
        // ---- amplitudes and probabilities of the m output modes (samplingHelperFunctions.cpp:308-339):
        //      rank 0 of the cluster holds the complete addends
        if (crank == 0) {
            // weighted addends of the DISTINCT occupied input modes, compacted (thread 0: N <= 40 entries)
            if (tid == 0) {
                const double sc = sh.K > 0 ? ldexp(1.0, -(sh.K - 1)) : 1.0;
                int nd = 0;
                for (int j = 0; j < N; j++) {
                    if (!sh.first[j]) continue;
                    const int col = sh.pidx[j];
                    const double cw = (double)sh.cur[col] * sc;
                    sh.wadd[nd] = make_double2(cw * sh.addend[j].x, cw * sh.addend[j].y);
                    sh.wcol[nd] = (unsigned char)col;
                    nd++;
                }
                sh.nd = nd;
            }
            __syncthreads();
            // all threads: output mode mm = tid % m, the distinct columns dealt round-robin over the
            // kLapThreads / m thread groups; partial amplitudes are combined through shared memory
            const int nd = sh.nd;
            const int groups = m > 64 ? 1 : (kLapThreads / m > 4 ? 4 : kLapThreads / m);
            double2* const amp = &sh.cred[0][0];  // [4][64] partial amplitudes; `cred` is free by now
            static_assert(kLapMaxCluster * kLapMaxPhotons >= 4 * 64, "amplitude scratch aliases cred");
            if (groups > 1) {
                const int mm = tid % m, grp = tid / m;
                double re = 0.0, im = 0.0;
                if (grp < groups) {
                    for (int k = grp; k < nd; k += groups) {
                        const double2 wv = sh.wadd[k];
                        const double2 u = a.U[(size_t)mm * a.ustride + sh.wcol[k]];
                        re += wv.x * u.x - wv.y * u.y;
                        im += wv.x * u.y + wv.y * u.x;
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
                    for (int k = 0; k < nd; k++) {
                        const double2 wv = sh.wadd[k];
                        const double2 u = a.U[(size_t)mm * a.ustride + sh.wcol[k]];
                        re += wv.x * u.x - wv.y * u.y;
                        im += wv.x * u.y + wv.y * u.x;
                    }
                    sh.p[mm] = re * re + im * im;
                }
            }
            __syncthreads();
            // probability_sum (sequential, as the reference forms it) and the normalised layer
            if (tid == 0) {
                double tot = 0.0;
                for (int i = 0; i < m; i++) tot += sh.p[i];
                sh.tot = tot;
            }
            __syncthreads();
            if (a.mode == 0) {
                const double tot = sh.tot;
                for (int i = tid; i < m; i += kLapThreads) sh.p[i] = sh.p[i] / tot;
                __syncthreads();
                if (tid == 0) {
                    // sample_from_pmf (samplingHelperFunctions.cpp:350-374): first index whose cumulative
                    // normalised probability reaches the draw, 0 if none does
                    const double r = a.rnd[(size_t)s * a.nmax + k];
                    int chosen = 0;
                    double acc = 0.0;
                    for (int i = 0; i < m; i++) {
                        acc += sh.p[i];
                        if (acc >= r) { chosen = i; break; }
                    }
                    for (unsigned int r2 = 0; r2 < csize; r2++) cluster.map_shared_rank(&sh, r2)->chosen = chosen;
                }
            }
        }
        if (csize > 1) cluster.sync(); else __syncthreads();
        if (a.mode == 0) {
            if (tid == 0) sh.sample[sh.chosen]++;
            __syncthreads();
        }
    }
    if (crank == 0) {
        if (a.mode == 0) {
            for (int i = tid; i < m; i += kLapThreads) a.samples_out[(size_t)s * m + i] = (long long)sh.sample[i];
        } else {
            const double tot = sh.tot;
            for (int i = tid; i < m; i += kLapThreads) a.pmf_out[(size_t)s * m + i] = sh.p[i] / tot;
        }
    }
}

}  // namespace pqb200
