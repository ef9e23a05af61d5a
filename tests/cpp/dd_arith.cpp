// dd_arith.cpp -- CPU check of csrc/dd.cuh, the double-double primitives behind every accumulation of the
// kernels and behind pqb200_permanent_dd (the engine's counterpart of the reference's precision mixing:
// long double / MPFR requests and common/source/PicComplexM.hpp:125-147).  Truth: __float128 (113 bits).
// Built and run by tests/test_host_logic_cpu.py with the CUDA qualifiers defined away.
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#define __host__
#define __device__
#define __forceinline__ inline
#include "../../piquassoboost_b200/csrc/dd.cuh"

using namespace pqb200;
typedef __float128 q;

static uint64_t g_state = 0x9e3779b97f4a7c15ull;
static uint64_t rnd64() {  // splitmix64
    uint64_t z = (g_state += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static double rnd_double(int spread) {  // full 53-bit mantissa, exponent in [-spread, spread], random sign
    const double m = 1.0 + (double)(rnd64() >> 11) * 0x1.0p-53;
    const int e = (int)(rnd64() % (uint64_t)(2 * spread + 1)) - spread;
    return ((rnd64() & 1) ? -m : m) * ldexp(1.0, e);
}
static dd rnd_dd(int spread) {  // normalised pair: |lo| <= ulp(hi) / 2
    dd x;
    x.hi = rnd_double(spread);
    x.lo = x.hi * 0x1.0p-54 * ((double)(int64_t)(rnd64() >> 12) * 0x1.0p-52 * 2.0 - 1.0) * 0.99;
    const double s = x.hi + x.lo;
    x.lo = x.lo - (s - x.hi);
    x.hi = s;
    return x;
}
static q val(const dd& x) { return (q)x.hi + (q)x.lo; }
static q qabs(q x) { return x < 0 ? -x : x; }
static double worst[8];
static void note(int k, q got, q want, q scale) {
    if (scale == 0) return;
    const double r = (double)(qabs(got - want) / scale);
    if (r > worst[k]) worst[k] = r;
}

int main() {
    const double kTiny = 0x1.0p-100;  // bound asserted for one operation (the algorithms give ~2^-104)
    for (int it = 0; it < 400000; it++) {
        const int spread = it % 3 == 0 ? 2 : it % 3 == 1 ? 40 : 300;
        // two_sum: error-free for every pair (no ordering assumption)
        const double a = rnd_double(spread), b = rnd_double(spread);
        double s, e;
        two_sum(a, b, s, e);
        if ((q)s + (q)e != (q)a + (q)b || s != a + b) { printf("FAIL two_sum %a %a\n", a, b); return 1; }
        const dd x = rnd_dd(spread > 200 ? 200 : spread), y = rnd_dd(spread > 200 ? 200 : spread);
        dd acc = x;
        dd_add_d(acc, a);
        note(0, val(acc), val(x) + (q)a, qabs(val(x)) + qabs((q)a));
        acc = x;
        dd_add_dd(acc, y);
        note(1, val(acc), val(x) + val(y), qabs(val(x)) + qabs(val(y)));
        const double f = rnd_double(spread > 200 ? 200 : spread);
        note(2, val(dd_mul_d(x, f)), val(x) * (q)f, qabs(val(x) * (q)f));
        note(3, val(dd_mul_dd(x, y)), val(x) * val(y), qabs(val(x) * val(y)));
        cdd u, v;
        u.re = x; u.im = y; v.re = rnd_dd(40); v.im = rnd_dd(40);
        const cdd w = cdd_mul(u, v);
        const q sc = (qabs(val(u.re)) + qabs(val(u.im))) * (qabs(val(v.re)) + qabs(val(v.im)));
        note(4, val(w.re), val(u.re) * val(v.re) - val(u.im) * val(v.im), sc);
        note(4, val(w.im), val(u.re) * val(v.im) + val(u.im) * val(v.re), sc);
        // results stay normalised: |lo| <= ulp(hi)
        if (fabs(acc.lo) > fabs(acc.hi) * 0x1.0p-52 || fabs(w.re.lo) > fabs(w.re.hi) * 0x1.0p-52) { printf("FAIL normalisation\n"); return 1; }
    }
    for (int k = 0; k < 5; k++)
        if (!(worst[k] < kTiny)) { printf("FAIL op %d rel err %.3g\n", k, worst[k]); return 1; }
    // a long cancelling sum, the kernels' pattern: terms 1e8 times the result, folded one by one
    dd acc = {0.0, 0.0};
    cdd cacc = cdd_zero();
    q truth = 0, mass = 0;
    double plain = 0.0;
    for (int i = 0; i < 2000000; i++) {
        const double t = rnd_double(30) * ((i & 1) ? 1.0 : -1.0);
        dd_add_d(acc, t);
        cdd c = cdd_zero();
        c.re.hi = t;
        c.im.hi = -t;
        cdd_add(cacc, c);
        truth += (q)t;
        mass += qabs((q)t);
        plain += t;
    }
    const double e_dd = (double)(qabs(val(acc) - truth) / mass), e_plain = (double)(qabs((q)plain - truth) / mass);
    const double e_c = (double)((qabs(val(cacc.re) - truth) + qabs(val(cacc.im) + truth)) / mass);
    if (!(e_dd < 0x1.0p-95) || !(e_c < 0x1.0p-95)) { printf("FAIL long sum %.3g %.3g\n", e_dd, e_c); return 1; }
    printf("ok worst rel err: add_d %.2g add_dd %.2g mul_d %.2g mul_dd %.2g cmul %.2g | 2e6-term sum: dd %.2g (plain double %.2g) of sum|x|\n",
           worst[0], worst[1], worst[2], worst[3], worst[4], e_dd, e_plain);
    return 0;
}
