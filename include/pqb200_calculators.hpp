// pqb200_calculators.hpp -- C++ calculator classes with the reference's names and call signatures,
// forwarding to the C ABI (include/pqb200.h).  This is the shim a piquassoboost maintainer drops in
// place of the TBB bodies: each class keeps `calculate(...) -> Complex16` and the reference's error
// convention (`throw std::string`).  It is templated on the container types so that it compiles both
// against the reference's own `pic::matrix` / `pic::PicState_int64`
// (common/source/matrix_base.hpp:41-561: members rows, cols, stride, get_data();
//  common/source/PicState_base.hpp:34-226: size(), get_data()) and against any struct with the same
// members (tests/cpp/shim_smoke.cpp).
//
//   reference class / method                                      file:line (piquassoboost/...)
//   BBFGPermanentCalculator::calculate(mtx, use_extended, use_inf) sampling/source/BBFGPermanentCalculator.cpp:89-126
//   GlynnPermanentCalculator<M,P>::calculate(mtx)                  sampling/source/GlynnPermanentCalculator.hpp:47-57
//   GlynnPermanentCalculatorSimple<M,P>::calculate(mtx)            sampling/source/GlynnPermanentCalculatorSimple.hpp:47-53
//   BBFGPermanentCalculatorRepeated::calculate(mtx, col, row, ..)  sampling/source/BBFGPermanentCalculatorRepeated.cpp:74-141
//   GlynnPermanentCalculatorRepeated<M,P>::calculate(mtx, in, out) sampling/source/GlynnPermanentCalculatorRepeated.hpp:49-74
//   BatchednPermanentCalculator::{add,calculate,reserve_space,clear} sampling/source/BatchedPermanentCalculator.cpp:81-100,109-296,304-325
//   compute_pmf(U, sample, current_input)                          sampling/simulation_strategies/source/samplingHelperFunctions.cpp:78-342
#pragma once
#include <complex>
#include <string>
#include <vector>

#include "pqb200.h"

namespace pqb200 {

// Inside the reference tree define PQB200_COMPLEX16 to pic::Complex16 before including this header
// (common/source/PicComplex.hpp:34-409: layout-compatible with std::complex<double>, constructible
// from (re, im)); stand-alone users get std::complex<double>.
#ifdef PQB200_COMPLEX16
using Complex16 = PQB200_COMPLEX16;
#else
using Complex16 = std::complex<double>;
#endif

inline void check(int rc) {
    if (rc != PQB200_OK) throw std::string(pqb200_last_error());
}

/// `ndev` handed to the C ABI by every shim: 0 = automatic (the library sizes the device count from the work
/// of the call: small permanents run on ONE GPU, an n=40 permanent on all of them; include/pqb200.h)
inline int& default_ndev() {
    static int n = 0;
    return n;
}

class BBFGPermanentCalculator {
public:
    template <class Matrix>
    Complex16 calculate(Matrix& mtx, bool use_extended = false, bool use_inf = false) {
        (void)use_extended;  // one fast numeric mode serves double and long double requests
        if (use_inf) {       // MPFR in the reference (BBFGPermanentCalculator.cpp:100-107): double-double here
            double dd[4];
            check(pqb200_permanent_dd(reinterpret_cast<const double*>(mtx.get_data()), mtx.rows, mtx.cols,
                                      mtx.stride, default_ndev(), dd));
            return Complex16(dd[0] + dd[1], dd[2] + dd[3]);
        }
        double out[2];
        check(pqb200_permanent(reinterpret_cast<const double*>(mtx.get_data()), mtx.rows, mtx.cols,
                               mtx.stride, default_ndev(), out));
        return Complex16(out[0], out[1]);
    }
};

class GlynnPermanentCalculator {
public:
    template <class Matrix>
    Complex16 calculate(Matrix& mtx) {
        return BBFGPermanentCalculator().calculate(mtx);
    }
};
using GlynnPermanentCalculatorDouble = GlynnPermanentCalculator;
using GlynnPermanentCalculatorLongDouble = GlynnPermanentCalculator;

class GlynnPermanentCalculatorSimple {
public:
    template <class Matrix>
    Complex16 calculate(Matrix& mtx) {
        if (mtx.rows != mtx.cols) return Complex16(0.0, 0.0);  // GlynnPermanentCalculatorSimple.hpp:67-70
        return BBFGPermanentCalculator().calculate(mtx);
    }
};

class BBFGPermanentCalculatorRepeated {
public:
    /// col_mult = input state, row_mult = output state (reference argument order).  Power form as in the
    /// reference (no photon-number cap); ill-conditioned inputs are re-evaluated in double-double by the
    /// library itself (pqb200_last_repeated_bound tells), use_inf forces that mode.
    template <class Matrix, class State>
    Complex16 calculate(Matrix& mtx, State& col_mult, State& row_mult, bool use_extended = false,
                        bool use_inf = false) {
        (void)use_extended;
        std::vector<int64_t> c(col_mult.size()), r(row_mult.size());
        for (size_t i = 0; i < c.size(); i++) c[i] = (int64_t)col_mult[i];
        for (size_t i = 0; i < r.size(); i++) r[i] = (int64_t)row_mult[i];
        if (use_inf) {  // MPFR in the reference (BBFGPermanentCalculatorRepeated.cpp:117-124): double-double here
            double dd[4];
            check(pqb200_permanent_repeated_dd(reinterpret_cast<const double*>(mtx.get_data()), mtx.rows,
                                               mtx.cols, mtx.stride, r.data(), c.data(), default_ndev(), dd));
            return Complex16(dd[0] + dd[1], dd[2] + dd[3]);
        }
        double out[2];
        check(pqb200_permanent_repeated(reinterpret_cast<const double*>(mtx.get_data()), mtx.rows,
                                        mtx.cols, mtx.stride, r.data(), c.data(), default_ndev(), out));
        return Complex16(out[0], out[1]);
    }
};

class GlynnPermanentCalculatorRepeated {
public:
    template <class Matrix, class State>
    Complex16 calculate(Matrix& mtx, State& input_state, State& output_state) {
        return BBFGPermanentCalculatorRepeated().calculate(mtx, input_state, output_state);
    }
};
using GlynnPermanentCalculatorRepeatedDouble = GlynnPermanentCalculatorRepeated;
using GlynnPermanentCalculatorRepeatedLongDouble = GlynnPermanentCalculatorRepeated;

/// Accumulates (input_state, output_state) pairs against one interferometer and evaluates them in ONE
/// batched call (BatchedPermanentCalculator.cpp:81-100,109-296; there: a CPU loop or DFE batches).
/// As in the reference the states are borrowed pointers that must outlive calculate(), the container
/// must be sized with reserve_space() before add(), and calculate() clears the list.
template <class Matrix, class State>
class BatchednPermanentCalculator {
public:
    BatchednPermanentCalculator() : mtx_(nullptr) {}
    explicit BatchednPermanentCalculator(Matrix& interferometer_matrix_in) : mtx_(&interferometer_matrix_in) {}

    void reserve_space(size_t num_of_permanents) {
        input_states_.resize(num_of_permanents, nullptr);
        output_states_.resize(num_of_permanents, nullptr);
    }

    void add(State* input_state, State* output_state, size_t index) {
        int64_t si = 0, so = 0;
        for (size_t i = 0; i < input_state->size(); i++) si += (int64_t)(*input_state)[i];
        for (size_t i = 0; i < output_state->size(); i++) so += (int64_t)(*output_state)[i];
        if (si != so)
            throw std::string("BatchednPermanentCalculator::add:  Number of input and output states should be equal");
        if (index >= input_states_.size() || index >= output_states_.size())
            throw std::string("BatchednPermanentCalculator::add:  Containers must be resized");
        input_states_[index] = input_state;
        output_states_[index] = output_state;
    }

    /// every supported `lib` id runs the CUDA engine; returns the permanents in insertion order
    std::vector<Complex16> calculate(int lib = 0) {
        (void)lib;
        if (!mtx_ || mtx_->rows == 0 || mtx_->cols == 0)
            throw std::string("BatchednPermanentCalculator::calculate:  interferometer matrix is zero");
        const size_t count = input_states_.size(), R = mtx_->rows, C = mtx_->cols;
        std::vector<int64_t> rm(count * R), cm(count * C);
        for (size_t k = 0; k < count; k++) {
            if (!input_states_[k] || !output_states_[k])
                throw std::string("BatchednPermanentCalculator::calculate:  unset state pair");
            for (size_t i = 0; i < R; i++) rm[k * R + i] = (int64_t)(*output_states_[k])[i];
            for (size_t j = 0; j < C; j++) cm[k * C + j] = (int64_t)(*input_states_[k])[j];
        }
        std::vector<double> out(2 * count);
        if (count)
            check(pqb200_permanent_repeated_batch(reinterpret_cast<const double*>(mtx_->get_data()), R, C,
                                                  mtx_->stride, rm.data(), cm.data(), count,
                                                  default_ndev(), out.data()));
        std::vector<Complex16> ret(count);
        for (size_t k = 0; k < count; k++) ret[k] = Complex16(out[2 * k], out[2 * k + 1]);
        clear();
        return ret;
    }

    void clear() {
        input_states_.clear();
        output_states_.clear();
    }

private:
    Matrix* mtx_;
    std::vector<State*> input_states_, output_states_;
};

/// compute_pmf for one shot; returns m probabilities
template <class Matrix, class State>
std::vector<double> compute_pmf(Matrix& interferometer_matrix, State& sample, State& current_input) {
    const size_t m = interferometer_matrix.rows;
    std::vector<int64_t> s(m), c(m);
    for (size_t i = 0; i < m; i++) {
        s[i] = (int64_t)sample[i];
        c[i] = (int64_t)current_input[i];
    }
    std::vector<double> pmf(m);
    check(pqb200_pmf_layer(reinterpret_cast<const double*>(interferometer_matrix.get_data()), m,
                           interferometer_matrix.stride, s.data(), c.data(), 1, default_ndev(),
                           pmf.data()));
    return pmf;
}

}  // namespace pqb200
