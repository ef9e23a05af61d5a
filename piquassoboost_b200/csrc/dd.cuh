// dd.cuh -- double-double ("dd") helpers used for every accumulation in the permanent kernels.
//
// The reference mixes precisions through a `precision_type` template parameter (double / x87 long
// double / MPFR) and, in the recursive Glynn calculators, a magnitude-bucketed accumulator
// (reference: piquassoboost/common/source/PicComplexM.hpp:125-147).  On the GPU there is no 80-bit
// type; the same goal -- do not lose small addends next to large partial sums -- is met by keeping
// every running sum as an unevaluated pair hi+lo (Knuth TwoSum), which gives ~106 mantissa bits,
// i.e. more than the reference's 64-bit long double accumulators.
#pragma once

namespace pqb200 {

struct dd {
    double hi, lo;
};

struct cdd {  // complex double-double
    dd re, im;
};

// error-free transformation: a + b = s + e exactly (no ordering assumption on |a|,|b|)
__host__ __device__ __forceinline__ void two_sum(double a, double b, double& s, double& e) {
    s = a + b;
    double bb = s - a;
    e = (a - (s - bb)) + (b - bb);
}

// acc += x   (x plain double)
__host__ __device__ __forceinline__ void dd_add_d(dd& acc, double x) {
    double s, e;
    two_sum(acc.hi, x, s, e);
    e += acc.lo;
    // renormalise (fast two-sum, |s| >= |e|)
    double hi = s + e;
    acc.lo = e - (hi - s);
    acc.hi = hi;
}

// acc += x   (x double-double)
__host__ __device__ __forceinline__ void dd_add_dd(dd& acc, const dd& x) {
    double s, e, t, f;
    two_sum(acc.hi, x.hi, s, e);
    two_sum(acc.lo, x.lo, t, f);
    e += t;
    double hi = s + e;
    e = e - (hi - s);
    e += f;
    double hi2 = hi + e;
    acc.lo = e - (hi2 - hi);
    acc.hi = hi2;
}

// x * f with f a plain double, error-free product of the leading parts (FMA two-prod)
__host__ __device__ __forceinline__ dd dd_mul_d(const dd& x, double f) {
    const double p = x.hi * f;
    const double e = fma(x.hi, f, -p) + x.lo * f;
    dd r;
    r.hi = p + e;
    r.lo = e - (r.hi - p);
    return r;
}

// x * y, both double-double (relative error ~ 2^-104)
__host__ __device__ __forceinline__ dd dd_mul_dd(const dd& x, const dd& y) {
    const double p = x.hi * y.hi;
    double e = fma(x.hi, y.hi, -p);
    e = fma(x.hi, y.lo, e);
    e = fma(x.lo, y.hi, e);
    dd r;
    r.hi = p + e;
    r.lo = e - (r.hi - p);
    return r;
}

__host__ __device__ __forceinline__ dd dd_neg(const dd& x) {
    dd r;
    r.hi = -x.hi;
    r.lo = -x.lo;
    return r;
}

__host__ __device__ __forceinline__ void cdd_add(cdd& acc, const cdd& x) {
    dd_add_dd(acc.re, x.re);
    dd_add_dd(acc.im, x.im);
}

// complex dd product
__host__ __device__ __forceinline__ cdd cdd_mul(const cdd& x, const cdd& y) {
    cdd r;
    r.re = dd_mul_dd(x.re, y.re);
    dd_add_dd(r.re, dd_neg(dd_mul_dd(x.im, y.im)));
    r.im = dd_mul_dd(x.re, y.im);
    dd_add_dd(r.im, dd_mul_dd(x.im, y.re));
    return r;
}

__host__ __device__ __forceinline__ cdd cdd_zero() {
    cdd z;
    z.re.hi = z.re.lo = z.im.hi = z.im.lo = 0.0;
    return z;
}

}  // namespace pqb200
