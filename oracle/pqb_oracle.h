/* oracle/pqb_oracle.h -- TEST INFRASTRUCTURE ONLY (never linked into or called by the product).
 *
 * A plain-C CPU restatement of the reference's permanent hot path, written for this repository
 * from the algorithm description in SURVEY.md Appendix A; every function cites the reference
 * file:line it follows.  It exists so that the CUDA path can be checked on a GPU box where
 * /root/reference is absent.  Pinned against (i) the reference's own golden vectors
 * (unittests/test_permanent.py:47-251) and (ii) the unmodified reference C++ compiled into
 * oracle/_ref/libpqbref.so -- see tests/test_oracle.py.
 *
 * prec: 0 = double, 1 = x87 long double (the reference's use_extended), 2 = __float128.
 * All functions return 0 on success, 1 when the photon numbers of the two states differ
 * (the reference throws std::string, BBFGPermanentCalculatorRepeated.hpp:119-124).
 */
#ifndef PQB_ORACLE_H
#define PQB_ORACLE_H
#include <stddef.h>
#include <stdint.h>

/* BBFGPermanentCalculator::calculate  (sampling/source/BBFGPermanentCalculator.hpp:87-235) */
int pqbo_bbfg(const double* a, size_t rows, size_t cols, int prec, double out[2]);

/* BBFGPermanentCalculatorRepeated::calculate (sampling/source/BBFGPermanentCalculatorRepeated.hpp:74-307)
 * col_mult = input state (cols entries), row_mult = output state (rows entries) */
int pqbo_bbfg_repeated(const double* a, size_t rows, size_t cols, const int64_t* col_mult,
                       const int64_t* row_mult, int prec, double out[2]);

/* adaptInterferometer (sampling/source/BatchedPermanentCalculator.cpp:412-451): gather rows with
 * out>0 and columns with in>0; out must hold m*m complex; returns the shape */
void pqbo_adapt_interferometer(const double* u, size_t m, const int64_t* input_state,
                               const int64_t* output_state, double* out, size_t* out_rows,
                               size_t* out_cols);

/* compute_pmf (sampling/simulation_strategies/source/samplingHelperFunctions.cpp:78-342) */
int pqbo_compute_pmf(const double* u, size_t m, const int64_t* sample, const int64_t* current_input,
                     int prec, double* pmf);

/* CGeneralizedCliffordsBSimulationStrategy::simulate with the shared std::mt19937
 * (CGeneralizedCliffordsBSimulationStrategy.cpp:137-316, samplingHelperFunctions.cpp:58-74,350-406);
 * samples_out: shots x m, C++ generation order */
int pqbo_simulate_B(const double* u, size_t m, uint64_t seed, const int64_t* input_state, int shots,
                    int64_t* samples_out);

/* std::mt19937 restated (for the RNG-order contract): fills out[n] with the first n
 * 32-bit outputs after seeding with (uint32_t)seed */
void pqbo_mt19937(uint32_t seed, uint32_t* out, size_t n);

/* worker threads used by the chunked loops (default: online cores, or $PQB_ORACLE_THREADS) */
int pqbo_threads(void);
void pqbo_set_threads(int n);

#endif
