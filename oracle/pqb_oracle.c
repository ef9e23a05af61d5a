/* oracle/pqb_oracle.c -- TEST INFRASTRUCTURE ONLY; see pqb_oracle.h. */
#include "pqb_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* ---- a tiny pthread parallel-for (the image's gcc wrapper has no libgomp) -------------------- */
static int g_threads = 0;
int pqbo_threads(void) {
    if (g_threads <= 0) {
        const char* e = getenv("PQB_ORACLE_THREADS");
        g_threads = e ? atoi(e) : (int)sysconf(_SC_NPROCESSORS_ONLN);
        if (g_threads <= 0) g_threads = 1;
    }
    return g_threads;
}
void pqbo_set_threads(int n) { g_threads = n > 0 ? n : 0; }

typedef void (*pqbo_job_fn)(void*, int64_t);
typedef struct { pqbo_job_fn fn; void* ctx; int64_t njobs; atomic_llong next; } pf_state;
static void* pf_worker(void* v) {
    pf_state* st = (pf_state*)v;
    for (;;) {
        long long j = atomic_fetch_add(&st->next, 1);
        if (j >= st->njobs) break;
        st->fn(st->ctx, (int64_t)j);
    }
    return NULL;
}
static void pqbo_parallel_for(int64_t njobs, pqbo_job_fn fn, void* ctx) {
    pf_state st;
    st.fn = fn; st.ctx = ctx; st.njobs = njobs; atomic_init(&st.next, 0);
    int nt = pqbo_threads();
    if ((int64_t)nt > njobs) nt = (int)njobs;
    if (nt <= 1) { pf_worker(&st); return; }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nt);
    for (int t = 1; t < nt; t++) pthread_create(&th[t], NULL, pf_worker, &st);
    pf_worker(&st);
    for (int t = 1; t < nt; t++) pthread_join(th[t], NULL);
    free(th);
}

#define REAL double
#define SFX _d
#include "pqb_oracle_impl.inc"
#undef REAL
#undef SFX

#define REAL long double
#define SFX _ld
#include "pqb_oracle_impl.inc"
#undef REAL
#undef SFX

#define REAL __float128
#define SFX _q
#include "pqb_oracle_impl.inc"
#undef REAL
#undef SFX

int pqbo_bbfg(const double* a, size_t rows, size_t cols, int prec, double out[2]) {
    /* edge cases of BBFGPermanentCalculator.hpp:90-95 */
    if (rows == 0 || cols == 0) { out[0] = 1; out[1] = 0; return 0; }
    if (rows >= cols + 2) { out[0] = 0; out[1] = 0; return 0; }
    if (prec == 0) bbfg_d(a, rows, cols, out);
    else if (prec == 1) bbfg_ld(a, rows, cols, out);
    else bbfg_q(a, rows, cols, out);
    return 0;
}

int pqbo_bbfg_repeated(const double* a, size_t rows, size_t cols, const int64_t* col_mult,
                       const int64_t* row_mult, int prec, double out[2]) {
    if (prec == 0) return bbfg_repeated_d(a, rows, cols, col_mult, row_mult, out);
    if (prec == 1) return bbfg_repeated_ld(a, rows, cols, col_mult, row_mult, out);
    return bbfg_repeated_q(a, rows, cols, col_mult, row_mult, out);
}

void pqbo_adapt_interferometer(const double* u, size_t m, const int64_t* input_state,
                               const int64_t* output_state, double* out, size_t* out_rows,
                               size_t* out_cols) {
    size_t nr = 0, nc = 0;
    for (size_t i = 0; i < m; i++) nr += output_state[i] > 0;
    for (size_t j = 0; j < m; j++) nc += input_state[j] > 0;
    size_t r = 0;
    for (size_t i = 0; i < m; i++) {
        if (output_state[i] <= 0) continue;
        size_t c = 0;
        for (size_t j = 0; j < m; j++) {
            if (input_state[j] <= 0) continue;
            out[2 * (r * nc + c)] = u[2 * (i * m + j)];
            out[2 * (r * nc + c) + 1] = u[2 * (i * m + j) + 1];
            c++;
        }
        r++;
    }
    *out_rows = nr;
    *out_cols = nc;
}

/* compute_pmf, samplingHelperFunctions.cpp:78-342: for every input mode i with photons,
 * addend[i] = perm(U[sample rows, (input - e_i) cols]) (cpp:137-172, always in double in the
 * reference, cpp:150-153; `prec` lets the tests ask for a long-double truth as well), then
 * pmf[k] = |sum_i input[i]*addend[i]*U[k,i]|^2, normalised (cpp:308-339). */
int pqbo_compute_pmf(const double* u, size_t m, const int64_t* sample, const int64_t* current_input,
                     int prec, double* pmf) {
    double* addr = (double*)calloc(2 * m, sizeof(double));
    double* sub = (double*)malloc(2 * m * m * sizeof(double));
    int64_t* in_loc = (int64_t*)malloc(m * sizeof(int64_t));
    int64_t* fin = (int64_t*)malloc(m * sizeof(int64_t));
    int64_t* fout = (int64_t*)malloc(m * sizeof(int64_t));
    for (size_t i = 0; i < m; i++) {
        if (current_input[i] <= 0) continue;
        memcpy(in_loc, current_input, m * sizeof(int64_t));
        in_loc[i]--;
        size_t nr, nc;
        pqbo_adapt_interferometer(u, m, in_loc, sample, sub, &nr, &nc);
        size_t a = 0, b = 0;
        for (size_t j = 0; j < m; j++) if (in_loc[j] > 0) fin[a++] = in_loc[j];
        for (size_t j = 0; j < m; j++) if (sample[j] > 0) fout[b++] = sample[j];
        double v[2];
        if (pqbo_bbfg_repeated(sub, nr, nc, fin, fout, prec, v)) { free(addr); free(sub); free(in_loc); free(fin); free(fout); return 1; }
        addr[2 * i] = v[0];
        addr[2 * i + 1] = v[1];
    }
    double tot = 0;
    for (size_t k = 0; k < m; k++) {
        double re = 0, im = 0;
        for (size_t i = 0; i < m; i++) {
            if (current_input[i] <= 0) continue;
            const double wr = (double)current_input[i] * addr[2 * i], wi = (double)current_input[i] * addr[2 * i + 1];
            const double ur = u[2 * (k * m + i)], ui = u[2 * (k * m + i) + 1];
            re += wr * ur - wi * ui;
            im += wr * ui + wi * ur;
        }
        pmf[k] = re * re + im * im;
        tot += pmf[k];
    }
    for (size_t k = 0; k < m; k++) pmf[k] /= tot;
    free(addr); free(sub); free(in_loc); free(fin); free(fout);
    return 0;
}

/* std::mt19937 (ISO C++ [rand.predef]: w=32 n=624 m=397 r=31 a=0x9908b0df u=11 d=0xffffffff s=7
 * b=0x9d2c5680 t=15 c=0xefc60000 l=18 f=1812433253), the engine behind rng_gen
 * (samplingHelperFunctions.cpp:58). */
typedef struct { uint32_t mt[624]; int idx; } mt_state;
static void mt_seed(mt_state* s, uint32_t seed) {
    s->mt[0] = seed;
    for (int i = 1; i < 624; i++) s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
    s->idx = 624;
}
static uint32_t mt_next(mt_state* s) {
    if (s->idx >= 624) {
        for (int i = 0; i < 624; i++) {
            uint32_t y = (s->mt[i] & 0x80000000u) | (s->mt[(i + 1) % 624] & 0x7fffffffu);
            s->mt[i] = s->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        s->idx = 0;
    }
    uint32_t y = s->mt[s->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}
void pqbo_mt19937(uint32_t seed, uint32_t* out, size_t n) {
    mt_state s;
    mt_seed(&s, seed);
    for (size_t i = 0; i < n; i++) out[i] = mt_next(&s);
}

/* CGeneralizedCliffordsBSimulationStrategy::simulate / fill_r_sample
 * (CGeneralizedCliffordsBSimulationStrategy.cpp:238-267, 286-316) with
 * update_input_by_single_photon (samplingHelperFunctions.cpp:377-406: pick the rng()%size-th
 * photon of the not-yet-used input photons), compute_pmf, sample_from_pmf (cpp:350-374: first
 * index whose cumulative sum >= rng()/rng.max(), default 0). */
int pqbo_simulate_B(const double* u, size_t m, uint64_t seed, const int64_t* input_state, int shots,
                    int64_t* samples_out) {
    mt_state rng;
    mt_seed(&rng, (uint32_t)seed);
    int64_t nphot = 0;
    for (size_t i = 0; i < m; i++) nphot += input_state[i];
    int64_t* cur = (int64_t*)malloc(m * sizeof(int64_t));
    int64_t* work = (int64_t*)malloc(m * sizeof(int64_t));
    double* pmf = (double*)malloc(m * sizeof(double));
    for (int s = 0; s < shots; s++) {
        int64_t* sample = samples_out + (size_t)s * m;
        memset(sample, 0, m * sizeof(int64_t));
        memset(cur, 0, m * sizeof(int64_t));
        memcpy(work, input_state, m * sizeof(int64_t)); /* photons still to be inserted */
        int64_t left = nphot;
        while (left > 0) {
            /* update_input_by_single_photon */
            uint32_t pick = mt_next(&rng) % (uint32_t)left;
            for (size_t i = 0; i < m; i++) {
                if ((int64_t)pick < work[i]) { work[i]--; cur[i]++; break; }
                pick -= (uint32_t)work[i];
            }
            left--;
            if (pqbo_compute_pmf(u, m, sample, cur, 0, pmf)) return 1;
            /* sample_from_pmf */
            const double r = (double)mt_next(&rng) / 4294967295.0;
            size_t chosen = 0;
            double acc = 0;
            for (size_t k = 0; k < m; k++) { acc += pmf[k]; if (acc >= r) { chosen = k; break; } }
            sample[chosen]++;
        }
    }
    free(cur); free(work); free(pmf);
    return 0;
}
