// oracle/ref_driver/ref_capi.cpp  --  TEST INFRASTRUCTURE ONLY (not part of the product).
//
// A C ABI over the UNMODIFIED reference calculator classes, so that the
// reference's own C++ (compiled from /root/reference by oracle/Makefile into
// oracle/_ref/libpqbref.so) can be driven from Python/ctypes by the tests, by
// tools/make_golden.py and by bench.py's CPU-baseline legs.  This file contains
// only glue written for this repository; no reference source is copied here.
//
// Reference entry points wrapped (paths relative to /root/reference/piquassoboost):
//   sampling/source/BBFGPermanentCalculator.cpp:89-126            BBFGPermanentCalculator::calculate
//   sampling/source/GlynnPermanentCalculator.hpp:47-57            GlynnPermanentCalculator<M,P>::calculate
//   sampling/source/GlynnPermanentCalculatorSimple.hpp:47-53      GlynnPermanentCalculatorSimple<M,P>::calculate
//   sampling/source/BBFGPermanentCalculatorRepeated.cpp:74-141    BBFGPermanentCalculatorRepeated::calculate
//   sampling/source/GlynnPermanentCalculatorRepeated.hpp:49-74    GlynnPermanentCalculatorRepeated<M,P>::calculate
//   sampling/source/CChinHuhPermanentCalculator.cpp               CChinHuhPermanentCalculator::calculate
//   sampling/source/BatchedPermanentCalculator.cpp:367-451        adaptInterferometer{,GlynnMultiplied}
//   sampling/simulation_strategies/source/samplingHelperFunctions.cpp:78-342   compute_pmf
//   sampling/simulation_strategies/source/CGeneralizedCliffordsBSimulationStrategy.cpp:137-278  simulate
//   sampling/simulation_strategies/source/CGeneralizedCliffordsBUniformLossesSimulationStrategy.cpp:110-251  simulate
//   sampling/simulation_strategies/source/CGeneralizedCliffordsBLossySimulationStrategy.cpp:146-330  simulate
//     (its cblas_zgemm / LAPACKE_zgesvd come from blas_shim.cpp: BLAS/LAPACK are not in this image)
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "BBFGPermanentCalculator.h"
#include "BBFGPermanentCalculatorRepeated.h"
#include "BatchedPermanentCalculator.h"
#include "CChinHuhPermanentCalculator.h"
#include "CGeneralizedCliffordsBLossySimulationStrategy.h"
#include "CGeneralizedCliffordsBSimulationStrategy.h"
#include "CGeneralizedCliffordsBUniformLossesSimulationStrategy.h"
#include "GlynnPermanentCalculator.h"
#include "GlynnPermanentCalculatorRepeated.h"
#include "GlynnPermanentCalculatorSimple.h"
#include "PicState.h"
#include "matrix.h"
#include "matrix_real.h"
#include "samplingHelperFunctions.h"

// BLAS==2 makes the reference call openblas_set_num_threads(1); provide the symbol.
extern "C" void openblas_set_num_threads(int) {}
extern "C" int openblas_get_num_threads(void) { return 1; }

namespace {

thread_local std::string g_error;

pic::matrix wrap_matrix(const double* a, size_t rows, size_t cols) {
    // borrowed view: matrix(data, rows, cols) does not take ownership
    return pic::matrix((pic::Complex16*)a, rows, cols);
}

pic::PicState_int64 make_state(const int64_t* s, size_t n) {
    pic::PicState_int64 st(n);
    int64_t total = 0;
    for (size_t i = 0; i < n; i++) {
        st[i] = s[i];
        total += s[i];
    }
    st.number_of_photons = total;
    return st;
}

template <typename F>
int guarded(F f) {
    try {
        f();
        return 0;
    } catch (std::string& e) {
        g_error = e;
    } catch (const char* e) {
        g_error = e;
    } catch (std::exception& e) {
        g_error = e.what();
    } catch (...) {
        g_error = "unknown exception";
    }
    return 1;
}

}  // namespace

extern "C" {

const char* ref_last_error(void) { return g_error.c_str(); }

int ref_hardware_concurrency(void) { return (int)std::thread::hardware_concurrency(); }

// precision: 0 = double, 1 = long double (use_extended)
int ref_bbfg(const double* a, size_t rows, size_t cols, int precision, double out[2]) {
    return guarded([&] {
        pic::matrix m = wrap_matrix(a, rows, cols);
        pic::BBFGPermanentCalculator calc;
        pic::Complex16 r = calc.calculate(m, precision != 0, false);
        out[0] = r.real();
        out[1] = r.imag();
    });
}

int ref_glynn(const double* a, size_t rows, size_t cols, int precision, double out[2]) {
    return guarded([&] {
        pic::matrix m = wrap_matrix(a, rows, cols);
        pic::Complex16 r;
        if (precision) {
            pic::GlynnPermanentCalculatorLongDouble calc;
            r = calc.calculate(m);
        } else {
            pic::GlynnPermanentCalculatorDouble calc;
            r = calc.calculate(m);
        }
        out[0] = r.real();
        out[1] = r.imag();
    });
}

int ref_glynn_simple(const double* a, size_t rows, size_t cols, int precision, double out[2]) {
    return guarded([&] {
        pic::matrix m = wrap_matrix(a, rows, cols);
        pic::Complex16 r;
        if (precision) {
            pic::GlynnPermanentCalculatorSimpleLongDouble calc;
            r = calc.calculate(m);
        } else {
            pic::GlynnPermanentCalculatorSimpleDouble calc;
            r = calc.calculate(m);
        }
        out[0] = r.real();
        out[1] = r.imag();
    });
}

// col_mult = input state, row_mult = output state (reference argument order)
int ref_bbfg_repeated(const double* a, size_t rows, size_t cols, const int64_t* col_mult,
                      const int64_t* row_mult, int precision, double out[2]) {
    return guarded([&] {
        pic::matrix m = wrap_matrix(a, rows, cols);
        pic::PicState_int64 cm = make_state(col_mult, cols);
        pic::PicState_int64 rm = make_state(row_mult, rows);
        pic::BBFGPermanentCalculatorRepeated calc;
        pic::Complex16 r = calc.calculate(m, cm, rm, precision != 0, false);
        out[0] = r.real();
        out[1] = r.imag();
    });
}

int ref_glynn_repeated(const double* a, size_t rows, size_t cols, const int64_t* input_state,
                       const int64_t* output_state, int precision, double out[2]) {
    return guarded([&] {
        pic::matrix m = wrap_matrix(a, rows, cols);
        pic::PicState_int64 in = make_state(input_state, cols);
        pic::PicState_int64 outs = make_state(output_state, rows);
        pic::Complex16 r;
        if (precision) {
            pic::GlynnPermanentCalculatorRepeatedLongDouble calc;
            r = calc.calculate(m, in, outs);
        } else {
            pic::GlynnPermanentCalculatorRepeatedDouble calc;
            r = calc.calculate(m, in, outs);
        }
        out[0] = r.real();
        out[1] = r.imag();
    });
}

int ref_chinhuh(const double* a, size_t rows, size_t cols, const int64_t* input_state,
                const int64_t* output_state, double out[2]) {
    return guarded([&] {
        pic::matrix m = wrap_matrix(a, rows, cols);
        pic::PicState_int64 in = make_state(input_state, cols);
        pic::PicState_int64 outs = make_state(output_state, rows);
        pic::CChinHuhPermanentCalculator calc;
        pic::Complex16 r = calc.calculate(m, in, outs);
        out[0] = r.real();
        out[1] = r.imag();
    });
}

// out must hold (#out>0) x (#in>0) complex numbers; the actual shape is returned
int ref_adapt_interferometer(const double* u, size_t m, const int64_t* input_state,
                             const int64_t* output_state, double* out, size_t* out_rows,
                             size_t* out_cols) {
    return guarded([&] {
        pic::matrix U = wrap_matrix(u, m, m);
        pic::PicState_int64 in = make_state(input_state, m);
        pic::PicState_int64 outs = make_state(output_state, m);
        pic::matrix r = pic::adaptInterferometer(U, in, outs);
        *out_rows = r.rows;
        *out_cols = r.cols;
        for (size_t i = 0; i < r.rows; i++)
            std::memcpy(out + 2 * i * r.cols, r.get_data() + i * r.stride, 16 * r.cols);
    });
}

// out must hold N x N complex numbers, N = sum(input_state)
int ref_adapt_interferometer_multiplied(const double* u, size_t m, const int64_t* input_state,
                                        const int64_t* output_state, double* out, size_t* out_n) {
    return guarded([&] {
        pic::matrix U = wrap_matrix(u, m, m);
        pic::PicState_int64 in = make_state(input_state, m);
        pic::PicState_int64 outs = make_state(output_state, m);
        pic::matrix r = pic::adaptInterferometerGlynnMultiplied(U, &in, &outs);
        *out_n = r.rows;
        for (size_t i = 0; i < r.rows; i++)
            std::memcpy(out + 2 * i * r.cols, r.get_data() + i * r.stride, 16 * r.cols);
    });
}

int ref_compute_pmf(const double* u, size_t m, const int64_t* sample, const int64_t* current_input,
                    double* pmf) {
    return guarded([&] {
        pic::matrix U = wrap_matrix(u, m, m);
        pic::PicState_int64 s = make_state(sample, m);
        pic::PicState_int64 in = make_state(current_input, m);
        pic::matrix_real p = pic::compute_pmf(U, s, in);
        for (size_t i = 0; i < m; i++) pmf[i] = p[i];
    });
}

// samples_out: shots x m int64, in C++ generation order
int ref_simulate_B(const double* u, size_t m, uint64_t seed, const int64_t* input_state, int shots,
                   int64_t* samples_out) {
    return guarded([&] {
        pic::matrix U = wrap_matrix(u, m, m);
        pic::CGeneralizedCliffordsBSimulationStrategy strategy(U, 0);
        strategy.seed(seed);
        pic::PicState_int64 in = make_state(input_state, m);
        std::vector<pic::PicState_int64> samples = strategy.simulate(in, shots);
        for (size_t k = 0; k < samples.size(); k++)
            for (size_t i = 0; i < m; i++) samples_out[k * m + i] = samples[k][i];
    });
}

// continues the reference's global generator (no reseed): a second simulate() of one strategy
int ref_simulate_B_two_calls(const double* u, size_t m, uint64_t seed, const int64_t* input_state,
                             int shots1, int shots2, int64_t* samples_out) {
    return guarded([&] {
        pic::matrix U = wrap_matrix(u, m, m);
        pic::CGeneralizedCliffordsBSimulationStrategy strategy(U, 0);
        strategy.seed(seed);
        pic::PicState_int64 in = make_state(input_state, m);
        size_t k = 0;
        for (int shots : {shots1, shots2}) {
            std::vector<pic::PicState_int64> samples = strategy.simulate(in, shots);
            for (size_t j = 0; j < samples.size(); j++, k++)
                for (size_t i = 0; i < m; i++) samples_out[k * m + i] = samples[j][i];
        }
    });
}

int ref_simulate_B_uniform_losses(const double* u, size_t m, uint64_t seed, const int64_t* input_state,
                                  int shots, double transmittance, int64_t* samples_out) {
    return guarded([&] {
        pic::matrix U = wrap_matrix(u, m, m);
        pic::CGeneralizedCliffordsBUniformLossesSimulationStrategy strategy(U, transmittance, 0);
        strategy.seed(seed);
        pic::PicState_int64 in = make_state(input_state, m);
        std::vector<pic::PicState_int64> samples = strategy.simulate(in, shots);
        for (size_t k = 0; k < samples.size(); k++)
            for (size_t i = 0; i < m; i++) samples_out[k * m + i] = samples[k][i];
    });
}

int ref_simulate_B_lossy(const double* u, size_t m, uint64_t seed, const int64_t* input_state, int shots,
                         int approximated_modes, int64_t* samples_out) {
    return guarded([&] {
        std::vector<double> copy(u, u + 2 * m * m);  // a real zgesvd overwrites its input
        pic::matrix U = wrap_matrix(copy.data(), m, m);
        pic::CGeneralizedCliffordsBLossySimulationStrategy strategy(U, approximated_modes, 0);
        strategy.seed(seed);
        pic::PicState_int64 in = make_state(input_state, m);
        std::vector<pic::PicState_int64> samples = strategy.simulate(in, shots);
        for (size_t k = 0; k < samples.size(); k++)
            for (size_t i = 0; i < m; i++) samples_out[k * m + i] = samples[k][i];
    });
}

}  // extern "C"
