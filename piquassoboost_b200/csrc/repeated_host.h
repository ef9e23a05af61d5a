// repeated_host.h -- host-side preparation of a repeated-row / repeated-column permanent for
// glynn_reppow_kernel and the double-double walk (dd_repeated.cu): the reference's filtering / anchor-row
// step, the orientation choice, the compact (power-form) matrix with the product side's columns sorted
// by multiplicity, the unit-modulus normalisation, and the per-chunk step table (toggled row, direction
// for both high parities, signed binomial weight).  Pure integer / table work; all floating-point
// summation happens on the GPU except the two `base` vectors (sums of at most D rows, formed in long
// double).
#pragma once
#include <stdint.h>

#include <math.h>

#include <algorithm>
#include <complex>
#include <string>
#include <vector>

#include "launch.h"

namespace pqb200 {

struct RepPrepared {
    bool trivial = false;  // result already in `value`
    double value[2] = {1.0, 0.0};
    bool plain_route = false;  // every Gray digit is binary and no product column repeats: plain kernel on `mat`
    int N = 0;                 // total photons (= expanded width)
    int R = 0;                 // rows of `mat` including the anchor row 0
    int C = 0;                 // distinct product-side columns
    std::vector<double> mat;   // R x C complex, interleaved, row-major (columns NOT expanded)
    std::vector<int> r;        // r[d] = multiplicity of row d+1, d = 0..R-2 (all >= 1)
    std::vector<int> cm;       // cm[j] = multiplicity of column j (all >= 1); the largest one last, the
                               // others in descending order (glynn_reppow.cuh)
    int photons = 0;
    double phase[4] = {1.0, 0.0, 0.0, 0.0};  // dd complex factor the kernel's sum is multiplied with
};

struct RepLimits {
    int max_rows;      // Gray-side rows besides the anchor
    int max_cols;      // distinct product-side columns
    int max_mult;      // largest multiplicity on either side
    int buf_bytes;     // parameter-bank bytes shared by the R x C matrix and the step tables
    int min_len;       // chunk steps that must still fit next to the matrix
    int max_plain_n;   // plain route: expanded width
};

// Reference semantics: BBFGPermanentCalculatorRepeated.cpp:74-141 and .hpp:74-140.
//   returns 0 ok, 3 photon-number mismatch, 2 bad argument, 4 shape beyond `lim` (message in err)
inline int rep_prepare(const double* a, size_t rows, size_t cols, size_t stride,
                       const int64_t* row_mult, const int64_t* col_mult, const RepLimits& lim,
                       RepPrepared& out, std::string& err) {
    int64_t sr = 0, sc = 0;
    for (size_t i = 0; i < rows; i++) {
        if (row_mult[i] < 0) { err = "negative row multiplicity"; return 2; }
        sr += row_mult[i];
    }
    for (size_t j = 0; j < cols; j++) {
        if (col_mult[j] < 0) { err = "negative column multiplicity"; return 2; }
        sc += col_mult[j];
    }
    if (sr != sc) {
        err = "BBFGPermanentCalculatorRepeated_Tasks::calculate:  Number of input and output states should be equal";
        return 3;
    }
    if (rows == 0 || cols == 0 || sr == 0) {  // hpp:126-128: empty permanent is 1
        out.trivial = true;
        return 0;
    }
    // keep rows / columns with non-zero multiplicity (adaptInterferometer, BatchedPermanentCalculator.cpp:412-451)
    std::vector<int> ri, ci;
    for (size_t i = 0; i < rows; i++) if (row_mult[i] > 0) ri.push_back((int)i);
    for (size_t j = 0; j < cols; j++) if (col_mult[j] > 0) ci.push_back((int)j);
    // orientation: perm(A) = perm(A^T); index the Gray space by the side with the smaller
    // prod(mult+1) (precedent: GlynnPermanentCalculatorRepeatedDFE.cpp:1138); ties keep rows.
    auto space = [](const std::vector<int>& idx, const int64_t* mult) {
        double s = 1.0;
        int64_t mn = INT64_MAX;
        for (int i : idx) { s *= (double)(mult[i] + 1); mn = std::min(mn, mult[i]); }
        return s * (double)mn / (double)(mn + 1);  // the anchor takes one copy of a min-multiplicity row
    };
    // does the orientation (Gray side g, product side p) fit the kernel's limits?
    auto fits = [&](const std::vector<int>& g, const int64_t* gmult, const std::vector<int>& p, const int64_t* pmult) {
        int64_t mn = INT64_MAX;
        for (int i : g) mn = std::min(mn, gmult[i]);
        int digits = (int)g.size() - (mn == 1 ? 1 : 0);
        for (int i : g) if (gmult[i] > lim.max_mult) return false;
        for (int j : p) if (pmult[j] > lim.max_mult) return false;
        if (digits > lim.max_rows || (int)p.size() > lim.max_cols) return false;
        const int C = (int)p.size(), CS = C % 16 == 0 ? C + 1 : C;
        return 16 * (digits + 1) * CS + 10 * lim.min_len + 64 <= lim.buf_bytes;
    };
    bool transpose = space(ci, col_mult) < space(ri, row_mult);
    const bool fit_pref = transpose ? fits(ci, col_mult, ri, row_mult) : fits(ri, row_mult, ci, col_mult);
    if (!fit_pref) {
        const bool fit_other = transpose ? fits(ri, row_mult, ci, col_mult) : fits(ci, col_mult, ri, row_mult);
        if (!fit_other) {
            err = "repeated permanent exceeds the limits of this build (at most " + std::to_string(lim.max_rows + 1) +
                  " distinct rows x " + std::to_string(lim.max_cols) + " distinct columns)";
            return 4;
        }
        transpose = !transpose;
    }
    const std::vector<int>& gi = transpose ? ci : ri;    // Gray side indices
    std::vector<int> pi = transpose ? ri : ci;           // product side indices
    const int64_t* gm = transpose ? col_mult : row_mult;
    const int64_t* pm = transpose ? row_mult : col_mult;
    auto elem = [&](int g, int p, int part) {
        const size_t i = transpose ? (size_t)p : (size_t)g, j = transpose ? (size_t)g : (size_t)p;
        return a[2 * (i * stride + j) + part];
    };
    // anchor: the Gray-side row with the smallest non-zero multiplicity (first one on ties, hpp:78)
    int amin = 0;
    for (int k = 1; k < (int)gi.size(); k++)
        if (gm[gi[k]] < gm[gi[amin]]) amin = k;
    std::vector<int> grow;  // Gray-side rows of `mat` after the anchor
    grow.push_back(gi[amin]);
    out.r.clear();
    for (int k = 0; k < (int)gi.size(); k++) {
        const int m = (int)gm[gi[k]] - (k == amin ? 1 : 0);
        if (m > 0) {
            grow.push_back(gi[k]);
            out.r.push_back(m);
        }
    }
    // product-side column order: the largest multiplicity last (its column is made real: a chain of
    // real multiplies), the others by descending multiplicity (stable)
    {
        int imax = 0;
        for (int k = 1; k < (int)pi.size(); k++)
            if (pm[pi[k]] >= pm[pi[imax]]) imax = k;
        const int last = pi[imax];
        pi.erase(pi.begin() + imax);
        std::stable_sort(pi.begin(), pi.end(), [&](int x, int y) { return pm[x] > pm[y]; });
        pi.push_back(last);
    }
    out.photons = (int)sr;
    out.N = (int)sr;
    out.R = (int)grow.size();
    out.C = (int)pi.size();
    out.cm.clear();
    for (int p : pi) out.cm.push_back((int)pm[p]);
    out.mat.assign((size_t)out.R * out.C * 2, 0.0);
    for (int i = 0; i < out.R; i++)
        for (int j = 0; j < out.C; j++) {
            out.mat[2 * ((size_t)i * out.C + j)] = elem(grow[i], pi[j], 0);
            out.mat[2 * ((size_t)i * out.C + j) + 1] = elem(grow[i], pi[j], 1);
        }
    if (out.R == 1) {  // single row: prod_j a_j^{c_j} (hpp:130-140)
        double re = 1.0, im = 0.0;
        for (int j = 0; j < out.C; j++) {
            const double xr = out.mat[2 * j], xi = out.mat[2 * j + 1];
            for (int k = 0; k < out.cm[j]; k++) {
                const double tr = re * xr - im * xi, ti = re * xi + im * xr;
                re = tr;
                im = ti;
            }
        }
        out.trivial = true;
        out.value[0] = re;
        out.value[1] = im;
        return 0;
    }
    out.plain_route = std::all_of(out.r.begin(), out.r.end(), [](int m) { return m == 1; }) &&
                      std::all_of(out.cm.begin(), out.cm.end(), [](int m) { return m == 1; });
    if (out.plain_route && out.N > lim.max_plain_n) {
        err = "matrix dimension exceeds " + std::to_string(lim.max_plain_n) +
              " (kernel-parameter resident matrix limit of this build)";
        return 4;
    }
    return 0;
}

// Unit-modulus row/column factors that make row 1 (the fastest digit's row) and the last column of
// `mat` real (glynn_reppow.cuh; same idea and numerics as plain_host.h).  Row i of `mat` stands for r_i
// copies (the anchor row 0 for one), column j for cm_j copies, so a factor enters the permanent with
// that power; out.phase receives the inverse of the total factor as a double-double complex number.
inline void rep_normalise(RepPrepared& out) {
    typedef std::complex<long double> cld;
    const int R = out.R, C = out.C;
    std::vector<cld> b((size_t)R * C);
    for (size_t k = 0; k < b.size(); k++) b[k] = cld(out.mat[2 * k], out.mat[2 * k + 1]);
    cld corr(1.0L, 0.0L);
    for (int j = 0; j < C; j++) {
        const cld v = b[(size_t)1 * C + j];
        const long double m = hypotl(v.real(), v.imag());
        if (m == 0.0L || v.imag() == 0.0L) continue;
        const cld f = std::conj(v) / m;
        for (int i = 0; i < R; i++) b[(size_t)i * C + j] *= f;
        b[(size_t)1 * C + j] = cld(m, 0.0L);
        const cld inv = std::conj(f) / std::norm(f);
        for (int k = 0; k < out.cm[j]; k++) corr *= inv;
    }
    for (int i = 0; i < R; i++) {
        if (i == 1) continue;
        const cld v = b[(size_t)i * C + (C - 1)];
        const long double m = hypotl(v.real(), v.imag());
        if (m == 0.0L || v.imag() == 0.0L) continue;
        const cld f = std::conj(v) / m;
        for (int j = 0; j < C; j++) b[(size_t)i * C + j] *= f;
        b[(size_t)i * C + (C - 1)] = cld(m, 0.0L);
        const cld inv = std::conj(f) / std::norm(f);
        const int mult = i == 0 ? 1 : out.r[i - 1];
        for (int k = 0; k < mult; k++) corr *= inv;
    }
    for (size_t k = 0; k < b.size(); k++) {
        out.mat[2 * k] = (double)b[k].real();
        out.mat[2 * k + 1] = (double)b[k].imag();
    }
    for (int i = 0; i < R; i++) out.mat[2 * ((size_t)i * C + (C - 1)) + 1] = 0.0;
    for (int j = 0; j < C; j++) out.mat[2 * ((size_t)1 * C + j) + 1] = 0.0;
    // |corr| == 1 up to rounding; renormalise so that long products of factors do not drift
    const long double cn = hypotl(corr.real(), corr.imag());
    if (cn > 0.0L) corr /= cn;
    const double rh = (double)corr.real(), ih = (double)corr.imag();
    out.phase[0] = rh;
    out.phase[1] = (double)(corr.real() - (long double)rh);
    out.phase[2] = ih;
    out.phase[3] = (double)(corr.imag() - (long double)ih);
}

struct RepStepHost {
    double w;
    int row;
    unsigned int flags;
};

struct RepPlan {
    int q = 0;
    unsigned int len = 1;
    unsigned long long nchunks = 1;
    std::vector<RepStepHost> table;
    std::vector<double> base;   // [2][C] complex interleaved
    std::vector<double> binom;  // [D][stride]
    int low_parity[2] = {0, 0};
    unsigned long long weight[64];
    unsigned char radix[64];
};

inline double binom_d(int r, int k) {
    if (k > r - k) k = r - k;
    long double c = 1.0L;  // every prefix is an integer: exact while it stays below 2^64, correctly rounded beyond
    for (int t = 1; t <= k; t++) c = c * (long double)(r - k + t) / (long double)t;
    return (double)c;
}

// C(r, k) as a double-double: exact through unsigned __int128 while the value stays below 2^120
// (the reference forms its binomials in __int128 as well, BBFGPermanentCalculatorRepeated.hpp:200,222,280)
inline void binom_dd(int r, int k, double& hi, double& lo) {
    if (k > r - k) k = r - k;
    unsigned __int128 c = 1;
    bool exact = true;
    for (int t = 1; t <= k && exact; t++) {
        if (c >> 112) exact = false;
        else c = c * (unsigned)(r - k + t) / (unsigned)t;
    }
    if (!exact) {
        hi = binom_d(r, k);
        lo = 0.0;
        return;
    }
    hi = (double)c;
    const unsigned __int128 h = (unsigned __int128)hi;
    lo = h <= c ? (double)(c - h) : -(double)(h - c);
}

// Choose the number of low digits and build the per-step table for them.
inline void rep_plan(const RepPrepared& pr, int max_resident_threads, unsigned long long shards,
                     unsigned int max_len, int binom_stride, RepPlan& pl) {
    const int D = pr.R - 1, N = pr.C;  // N: columns the kernel keeps sums for
    for (int d = 0; d < D; d++) pl.radix[d] = (unsigned char)(pr.r[d] + 1);
    double total = 1.0;
    for (int d = 0; d < D; d++) total *= pl.radix[d];
    // cost model: waves * (seed + len * step)
    double best = 1e300;
    int bestq = 0;
    unsigned long long len = 1;
    for (int q = 0; q <= D; q++) {
        if (q > 0) len *= pl.radix[q - 1];
        if (len > max_len) break;
        const double chunks = total / (double)len;
        const double per_shard = std::max(1.0, chunks / (double)shards);
        const double threads = std::min((double)max_resident_threads, per_shard);
        const double waves = (double)(unsigned long long)((per_shard + threads - 1) / threads);
        const double cost = waves * ((D - q) * (2.0 * N + 60.0) + 2.0 * N + (double)len * (2.0 * N + 4.0 * pr.photons + 14.0));
        if (cost < best) {
            best = cost;
            bestq = q;
        }
    }
    pl.q = bestq;
    pl.len = 1;
    for (int d = 0; d < pl.q; d++) pl.len *= pl.radix[d];
    pl.nchunks = 1;
    for (int d = pl.q; d < D; d++) {
        pl.weight[d] = pl.nchunks;
        pl.nchunks *= pl.radix[d];
    }
    for (int d = 0; d < pl.q; d++) pl.weight[d] = 0;
    // binomial table
    pl.binom.assign((size_t)std::max(D, 1) * binom_stride, 0.0);
    for (int d = 0; d < D; d++)
        for (int k = 0; k <= pr.r[d]; k++) pl.binom[(size_t)d * binom_stride + k] = binom_d(pr.r[d], k);
    // low digits: simulate the reflected mixed-radix counter for both high parities
    const int q = pl.q;
    std::vector<int> cnt(q, 0), g0(q, 0), g1(q, 0), n0(q), n1(q);
    auto gray = [&](int P, std::vector<int>& g) {
        int par = P;
        for (int d = q - 1; d >= 0; d--) {
            g[d] = par ? pr.r[d] - cnt[d] : cnt[d];
            par ^= g[d] & 1;
        }
    };
    gray(0, g0);
    gray(1, g1);
    pl.base.assign((size_t)2 * N * 2, 0.0);
    for (int P = 0; P < 2; P++) {
        const std::vector<int>& g = P ? g1 : g0;
        int ks = 0;
        for (int d = 0; d < q; d++) ks += g[d];
        pl.low_parity[P] = ks & 1;
        for (int j = 0; j < N; j++) {
            long double re = pr.mat[2 * j], im = pr.mat[2 * j + 1];
            for (int d = 0; d < q; d++) {
                const long double coef = (long double)(pr.r[d] - 2 * g[d]);
                re += coef * (long double)pr.mat[2 * ((size_t)(d + 1) * N + j)];
                im += coef * (long double)pr.mat[2 * ((size_t)(d + 1) * N + j) + 1];
            }
            pl.base[2 * ((size_t)P * N + j)] = (double)re;
            pl.base[2 * ((size_t)P * N + j) + 1] = (double)im;
        }
    }
    pl.table.assign(pl.len, RepStepHost{1.0, 0, 0u});
    auto weight_of = [&](const std::vector<int>& g) {
        double w = 1.0;
        for (int d = 0; d < q; d++) w *= pl.binom[(size_t)d * binom_stride + g[d]];
        return w;
    };
    pl.table[0].w = weight_of(g0);
    for (unsigned int t = 1; t < pl.len; t++) {
        int d = 0;
        for (;;) {  // n_aryGrayCodeCounter::next, n_aryGrayCodeCounter.cpp:203-216
            if (cnt[d] < pr.r[d]) { cnt[d]++; break; }
            cnt[d] = 0;
            d++;
        }
        gray(0, n0);
        gray(1, n1);
        int ch = -1;
        for (int e = 0; e < q; e++)
            if (n0[e] != g0[e]) ch = e;
        unsigned int flags = 0;
        if (n0[ch] > g0[ch]) flags |= 1u;
        if (n1[ch] > g1[ch]) flags |= 2u;
        g0 = n0;
        g1 = n1;
        pl.table[t].row = ch + 1;
        pl.table[t].flags = flags;
        pl.table[t].w = ((t & 1u) ? -1.0 : 1.0) * weight_of(g0);
    }
}

}  // namespace pqb200
