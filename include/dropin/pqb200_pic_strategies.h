// pqb200_pic_strategies.h -- the reference's three Clifford-Clifford "B" simulation strategy classes,
// under the reference's own names in namespace pic, with simulate() served by libpqb200.so (the whole
// photon loop of all shots on the GPU, include/pqb200.h: pqb200_simulate_B_next /
// _uniform_losses / _lossy).  Companion of pqb200_pic_calculators.h: with this directory in front of
// piquassoboost/sampling/simulation_strategies/source on the include path the reference's CPython
// strategy wrappers (simulation_strategies/GeneralizedCliffordsB*SimulationStrategy_wrapper.cpp) compile
// UNMODIFIED against these classes (oracle/Makefile: `make dropin`).
//
//   reference class                                              (piquassoboost/sampling/simulation_strategies/source/)
//   pic::CGeneralizedCliffordsBSimulationStrategy                CGeneralizedCliffordsBSimulationStrategy.h:40-127, .cpp:61-231
//   pic::CGeneralizedCliffordsBUniformLossesSimulationStrategy   CGeneralizedCliffordsBUniformLossesSimulationStrategy.h, .cpp:20-330
//   pic::CGeneralizedCliffordsBLossySimulationStrategy           CGeneralizedCliffordsBLossySimulationStrategy.h, .cpp:60-508
//
// Kept from the reference: the constructors seed the process-global generator with time(NULL)
// (.cpp:75), seed() forwards to it, Update_interferometer_matrix stores the matrix HEADER (the data stays
// the caller's), simulate() returns the samples in generation order; errors are thrown as std::string.
#ifndef PQB200_PIC_STRATEGIES_H
#define PQB200_PIC_STRATEGIES_H

#include <time.h>

#include <vector>

#include "PicState.h"
#include "matrix.h"

#ifndef PQB200_COMPLEX16
#define PQB200_COMPLEX16 pic::Complex16
#endif
#include "../pqb200_calculators.hpp"

// `lib` ids of the strategies' permanent back end (the reference declares them in
// sampling/source/BatchedPermanentCalculator.h:16-21, which its strategy headers include); every id is
// served by the one CUDA engine
#ifndef GlynnRep
#define GlynnRep 0
#define ChinHuh 1
#define GlynnRepSingleDFE 2
#define GlynnRepDualDFE 3
#define GlynnRepMultiSingleDFE 4
#define GlynnRepMultiDualDFE 5
#endif

namespace pic {

namespace pqb200_detail {
inline std::vector<PicState_int64> to_states(const std::vector<int64_t>& flat, size_t shots, size_t m) {
    std::vector<PicState_int64> samples;
    samples.reserve(shots);
    for (size_t s = 0; s < shots; s++) {
        PicState_int64 st(m);
        int64_t n = 0;
        for (size_t i = 0; i < m; i++) {
            st[i] = flat[s * m + i];
            n += st[i];
        }
        st.number_of_photons = n;
        samples.push_back(st);
    }
    return samples;
}
}  // namespace pqb200_detail

class CGeneralizedCliffordsBSimulationStrategy {
protected:
    matrix interferometer_matrix;
    int lib;

public:
    CGeneralizedCliffordsBSimulationStrategy() : lib(0) { seed(time(NULL)); }
    CGeneralizedCliffordsBSimulationStrategy(matrix& interferometer_matrix_in, int lib_in) : lib(lib_in) {
        Update_interferometer_matrix(interferometer_matrix_in);
        seed(time(NULL));
    }
    void seed(unsigned long long int value) { pqb200::check(pqb200_seed(value)); }
    void Update_interferometer_matrix(matrix& interferometer_matrix_in) { interferometer_matrix = interferometer_matrix_in; }
    std::vector<PicState_int64> simulate(PicState_int64& input_state_in, int samples_number) {
        const size_t m = interferometer_matrix.rows, shots = samples_number > 0 ? (size_t)samples_number : 0;
        if (input_state_in.size() != m) throw std::string("input state length does not match the interferometer");
        std::vector<int64_t> in(m), flat(shots * m);
        for (size_t i = 0; i < m; i++) in[i] = input_state_in[i];
        pqb200::check(pqb200_simulate_B_next(reinterpret_cast<const double*>(interferometer_matrix.get_data()), m,
                                            interferometer_matrix.stride, in.data(), shots, pqb200::default_ndev(),
                                            flat.data()));
        return pqb200_detail::to_states(flat, shots, m);
    }
};

class CGeneralizedCliffordsBUniformLossesSimulationStrategy {
protected:
    matrix interferometer_matrix;
    double transmittance;
    int lib;

public:
    CGeneralizedCliffordsBUniformLossesSimulationStrategy() : transmittance(1.0), lib(0) { seed(time(NULL)); }
    CGeneralizedCliffordsBUniformLossesSimulationStrategy(matrix& interferometer_matrix_in, double transmittance_in, int lib_in)
        : transmittance(transmittance_in), lib(lib_in) {
        Update_interferometer_matrix(interferometer_matrix_in);
        seed(time(NULL));
    }
    void seed(unsigned long long int value) { pqb200::check(pqb200_seed(value)); }
    void Update_interferometer_matrix(matrix& interferometer_matrix_in) { interferometer_matrix = interferometer_matrix_in; }
    std::vector<PicState_int64> simulate(PicState_int64& input_state_in, int samples_number) {
        const size_t m = interferometer_matrix.rows, shots = samples_number > 0 ? (size_t)samples_number : 0;
        if (input_state_in.size() != m) throw std::string("input state length does not match the interferometer");
        std::vector<int64_t> in(m), flat(shots * m);
        for (size_t i = 0; i < m; i++) in[i] = input_state_in[i];
        pqb200::check(pqb200_simulate_B_uniform_losses(reinterpret_cast<const double*>(interferometer_matrix.get_data()), m,
                                                      interferometer_matrix.stride, in.data(), shots, transmittance,
                                                      pqb200::default_ndev(), flat.data()));
        return pqb200_detail::to_states(flat, shots, m);
    }
};

class CGeneralizedCliffordsBLossySimulationStrategy {
protected:
    matrix interferometer_matrix;
    int number_of_approximated_modes;
    int lib;

public:
    CGeneralizedCliffordsBLossySimulationStrategy() : number_of_approximated_modes(0), lib(0) { seed(time(NULL)); }
    CGeneralizedCliffordsBLossySimulationStrategy(matrix& interferometer_matrix_in, int lib_in)
        : number_of_approximated_modes(0), lib(lib_in) {
        Update_interferometer_matrix(interferometer_matrix_in);
        seed(time(NULL));
    }
    CGeneralizedCliffordsBLossySimulationStrategy(matrix& interferometer_matrix_in, int number_of_approximated_modes_in, int lib_in)
        : number_of_approximated_modes(number_of_approximated_modes_in), lib(lib_in) {
        Update_interferometer_matrix(interferometer_matrix_in);
        seed(time(NULL));
    }
    void seed(unsigned long long int value) { pqb200::check(pqb200_seed(value)); }
    void set_approximated_modes_number(int value) { number_of_approximated_modes = value; }
    void Update_interferometer_matrix(matrix& interferometer_matrix_in) { interferometer_matrix = interferometer_matrix_in; }
    /// As in the reference, simulate() extracts the uniform losses from the stored matrix and KEEPS the
    /// rescaled matrix (CGeneralizedCliffordsBLossySimulationStrategy.cpp:386-452): the data behind the
    /// header (the caller's numpy array) is updated in place.
    std::vector<PicState_int64> simulate(PicState_int64& input_state_in, int samples_number) {
        const size_t m = interferometer_matrix.rows, shots = samples_number > 0 ? (size_t)samples_number : 0;
        if (input_state_in.size() != m) throw std::string("input state length does not match the interferometer");
        std::vector<int64_t> in(m), flat(shots * m);
        for (size_t i = 0; i < m; i++) in[i] = input_state_in[i];
        double uniform_loss = 0.0;
        pqb200::check(pqb200_simulate_B_lossy(reinterpret_cast<double*>(interferometer_matrix.get_data()), m,
                                             interferometer_matrix.stride, in.data(), shots, number_of_approximated_modes,
                                             pqb200::default_ndev(), &uniform_loss, flat.data()));
        return pqb200_detail::to_states(flat, shots, m);
    }
};

}  // namespace pic

#endif
