// Drop-in check against the REFERENCE's own container types: include/pqb200_calculators.hpp is
// instantiated with pic::matrix / pic::PicState_int64 (common/source/matrix.h, PicState.h of the
// reference tree) exactly as a maintainer's patched calculators would call it.  Compiled only where
// /root/reference exists (the dev container); the reference sources are used in place, never copied.
// Without a GPU the calls must throw std::string (the reference's error convention): exit code 10.
#include <cstdio>
#include <string>

#include "matrix.h"
#include "PicState.h"
#define PQB200_COMPLEX16 pic::Complex16
#include "pqb200_calculators.hpp"

int main() {
    const double s = 0.70710678118654752440;
    pic::matrix m(2, 2);
    m[0] = pic::Complex16(s, 0);
    m[1] = pic::Complex16(0, s);
    m[2] = pic::Complex16(s, 0);
    m[3] = pic::Complex16(0, -s);
    pic::PicState_int64 in(2), out(2);
    in[0] = 0; in[1] = 2;
    out[0] = 2; out[1] = 0;
    try {
        pic::Complex16 v = pqb200::BBFGPermanentCalculatorRepeated().calculate(m, in, out);
        pic::Complex16 w = pqb200::BBFGPermanentCalculator().calculate(m, /*use_extended=*/true);
        pic::Complex16 g = pqb200::GlynnPermanentCalculatorLongDouble().calculate(m);
        pqb200::BatchednPermanentCalculator<pic::matrix, pic::PicState_int64> batch(m);
        batch.reserve_space(1);
        batch.add(&in, &out, 0);
        std::vector<pqb200::Complex16> b = batch.calculate();
        pic::PicState_int64 sample(2);  // one photon placed so far; compute_pmf adds the second
        sample[0] = 1; sample[1] = 0;
        std::vector<double> pmf = pqb200::compute_pmf(m, sample, in);
        std::printf("repeated=(%.15g,%.15g) plain=(%.15g,%.15g) glynn=(%.15g,%.15g) batch=(%.15g,%.15g) pmf0=%.6g\n",
                    v.real(), v.imag(), w.real(), w.imag(), g.real(), g.imag(), b[0].real(), b[0].imag(), pmf[0]);
        if (std::abs(v - pic::Complex16(-1, 0)) > 1e-14 || std::abs(b[0] - v) > 1e-14 || std::abs(w - g) > 1e-15)
            return 2;
        return 0;
    } catch (std::string& e) {
        std::printf("threw std::string: %s\n", e.c_str());
        return 10;
    }
}
