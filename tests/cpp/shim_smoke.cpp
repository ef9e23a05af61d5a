// Compile-and-link check of include/pqb200_calculators.hpp against libpqb200.so with stand-in
// container types that have the reference containers' members.  With a GPU it also runs the
// reference's 2x2 known answer (unittests/test_permanent.py:143-149); without one it checks that the
// shim throws std::string (no CPU fallback).
#include <cstdio>
#include <complex>
#include <string>
#include <vector>

#include "pqb200_calculators.hpp"

struct Mat {
    size_t rows, cols, stride;
    std::vector<std::complex<double>> d;
    std::complex<double>* get_data() { return d.data(); }
};
struct State {
    std::vector<int64_t> v;
    size_t size() const { return v.size(); }
    int64_t operator[](size_t i) const { return v[i]; }
};

int main() {
    const double s = 0.70710678118654752440;
    Mat m{2, 2, 2, {{s, 0}, {0, s}, {s, 0}, {0, -s}}};
    State in{{0, 2}}, out{{2, 0}};
    try {
        auto v = pqb200::BBFGPermanentCalculatorRepeated().calculate(m, in, out);
        auto w = pqb200::GlynnPermanentCalculator().calculate(m);
        std::printf("repeated=(%.15g,%.15g) plain=(%.15g,%.15g)\n", v.real(), v.imag(), w.real(), w.imag());
        if (std::abs(v - std::complex<double>(-1, 0)) > 1e-14) return 2;
        // use_inf requests (MPFR in the reference) are served by the double-double modes
        auto vi = pqb200::BBFGPermanentCalculatorRepeated().calculate(m, in, out, false, true);
        auto wi = pqb200::BBFGPermanentCalculator().calculate(m, false, true);
        if (std::abs(vi - v) > 1e-15 || std::abs(wi - w) > 1e-15) return 5;
        // BatchednPermanentCalculator: two state pairs against the same interferometer
        State in2{{1, 1}}, out2{{1, 1}};
        pqb200::BatchednPermanentCalculator<Mat, State> batch(m);
        batch.reserve_space(2);
        batch.add(&in, &out, 0);
        batch.add(&in2, &out2, 1);
        auto b = batch.calculate();
        std::printf("batch=(%.15g,%.15g) (%.15g,%.15g)\n", b[0].real(), b[0].imag(), b[1].real(), b[1].imag());
        if (std::abs(b[0] - v) > 1e-14 || std::abs(b[1] - w) > 1e-14) return 3;
        bool threw = false;
        try {
            State bad{{1, 0}};
            batch.reserve_space(1);
            batch.add(&bad, &out, 0);
        } catch (std::string&) {
            threw = true;
        }
        if (!threw) return 4;
        return 0;
    } catch (std::string& e) {
        std::printf("threw std::string: %s\n", e.c_str());
        return 10;
    }
}
