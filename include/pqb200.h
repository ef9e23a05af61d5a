/* pqb200.h -- C ABI of libpqb200.so, the B200 (sm_100a) permanent engine behind piquassoboost's
 * permanent calculators.
 *
 * This is the drop-in boundary: plain pointers and sizes only, no C++/torch/numpy types.  Every
 * entry point names the reference interface it replaces (paths relative to the reference tree,
 * piquassoboost/...).  The precedent for an accelerator C ABI in the reference is the FPGA ("DFE")
 * back end: init / compute(batch count, caller-allocated output) / release, resolved with dlsym
 * (sampling/source/GlynnPermanentCalculatorDFE.cpp:47-58,117-145).
 *
 * Conventions
 *   - matrices are row-major complex128 given as interleaved (re,im) doubles; `stride` is the
 *     distance between rows in complex elements (common/source/matrix_base.hpp:44-51);
 *   - all inputs are borrowed for the duration of the call only, outputs are caller allocated;
 *   - every function returns 0 on success; on failure a non-zero code, and pqb200_last_error()
 *     returns a message for the calling thread.  The C++ shims turn that into `throw std::string`
 *     (the reference's error convention, e.g. BBFGPermanentCalculatorRepeated.hpp:122-123);
 *   - there is NO CPU fallback: without a usable CUDA device every compute call fails;
 *   - the library is re-entrant: every call owns a WORK SLOT (stream + workspaces; two per device)
 *     from its first enqueue to its last synchronise, calls on different GPUs -- and two calls on the
 *     same GPU -- run concurrently, so it may be called from the TBB workers of
 *     samplingHelperFunctions.cpp:137 (the batched entry points remain the intended replacement for
 *     that loop).  The caller's current CUDA device is restored before every return.
 *   - `ndev`: GPUs one call is sharded over.
 *       0  automatic: as many devices as the call's work is worth (about one per millisecond of
 *          estimated kernel time, at most pqb200_init's count) -- a 14x14 permanent runs on ONE GPU,
 *          an n=40 permanent on all of them.  A one-device call runs on the calling thread's current
 *          CUDA device (cudaSetDevice / torch.cuda.set_device; device 0 if never set) and moves to
 *          another initialised GPU when that device's slots are busy;
 *       1  the calling thread's current CUDA device;
 *      >=2 devices 0 .. ndev-1.
 *   - asynchronous entry points (pqb200_permanent_shard with d_out, pqb200_permanent_nccl_async) leave
 *     work in flight on the caller's stream; the library orders every later use of the same slot
 *     behind it (cudaStreamWaitEvent), so they may be mixed freely with the blocking calls.
 */
#ifndef PQB200_H
#define PQB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PQB200_OK 0
#define PQB200_ERR_CUDA 1      /* CUDA runtime / driver / NCCL failure */
#define PQB200_ERR_ARG 2       /* invalid argument (shape, size limit, null pointer) */
#define PQB200_ERR_MISMATCH 3  /* sum(row multiplicities) != sum(column multiplicities) */

/* Library life cycle.  Replaces initialize_DFE / releive_DFE (GlynnPermanentCalculatorDFE.cpp:125-166).
 * pqb200_init(ndev): bind the first `ndev` visible CUDA devices (0 = all); idempotent; compute calls
 * initialise lazily with all devices if it was never called. */
int pqb200_init(int ndev);
void pqb200_shutdown(void);
int pqb200_device_count(void);
const char* pqb200_last_error(void);
const char* pqb200_version(void);

/* perm(A) for one matrix.  Replaces
 *   BBFGPermanentCalculator::calculate(matrix&, use_extended, use_inf)   sampling/source/BBFGPermanentCalculator.cpp:89-126
 *   GlynnPermanentCalculator<M,P>::calculate(matrix&)                    sampling/source/GlynnPermanentCalculator.hpp:47-57
 *   GlynnPermanentCalculatorSimple<M,P>::calculate(matrix&)              sampling/source/GlynnPermanentCalculatorSimple.hpp:47-53
 * Edge semantics of the reference are kept: rows==0 || cols==0 -> 1; rows >= cols+2 -> 0; otherwise
 * 2^-(rows-1) * sum over the 2^(rows-1) delta vectors of prod over the `cols` column sums. */
int pqb200_permanent(const double* a, size_t rows, size_t cols, size_t stride, int ndev,
                     double out[2]);

/* One shard of perm(A): the partial, UNSCALED double-double sum over Gray-code chunk range
 * shard_index of shard_count, computed on CUDA device `device`.  This is what one rank of a
 * one-process-per-GPU job calls; the 4 doubles per rank (re.hi re.lo im.hi im.lo) are then
 * all-gathered (NCCL) and folded with pqb200_permanent_combine.  When d_out is non-NULL it must be a
 * device pointer on `device`: the partial is left there (no host sync) so that the collective can
 * be enqueued directly behind the kernel on `stream` (a cudaStream_t, NULL = the library's own). */
int pqb200_permanent_shard(const double* a, size_t rows, size_t cols, size_t stride, int device,
                           uint64_t shard_index, uint64_t shard_count, double out_dd[4],
                           double* d_out, void* stream);
/* Fold `count` partials (4 doubles each) in index order with double-double adds and apply the
 * exact 2^-(rows-1) scaling (reference: BBFGPermanentCalculator.hpp:227-232). */
int pqb200_permanent_combine(const double* partials_dd, size_t count, size_t rows, double out[2]);

/* ---- one process per GPU, collective inside the library (NCCL resolved with dlopen at first use) ----
 * the Gray-code space of ONE permanent is sharded over the ranks and ONE
 * ncclAllGather of 4 doubles per rank, enqueued on the compute stream directly behind the kernel,
 * combines the partials; every rank folds them in rank order (double-double) and gets the same bits.
 * The reference has no counterpart (its MPI mode shards shots:
 * sampling/simulation_strategies/source/CGeneralizedCliffordsBSimulationStrategy.cpp:158-231).
 *   rank 0: pqb200_nccl_unique_id(id); broadcast the 128 bytes with the job's launcher (MPI, torch, file);
 *   every rank: pqb200_nccl_init(local_device, rank, world, id); then pqb200_permanent_nccl(...) collectively. */
int pqb200_nccl_unique_id(char id[128]);
int pqb200_nccl_init(int device, int rank, int world, const char id[128]);
int pqb200_nccl_shutdown(void);
int pqb200_permanent_nccl(const double* a, size_t rows, size_t cols, size_t stride, double out[2]);
/* Non-blocking form: kernel + all-gather enqueued on `stream` (a cudaStream_t; NULL = the library's own);
 * d_gathered (device, 4*world doubles; NULL = the library's buffer) receives the partials in rank order,
 * to be folded with pqb200_permanent_combine after the caller synchronised the stream. */
int pqb200_permanent_nccl_async(const double* a, size_t rows, size_t cols, size_t stride,
                                double* d_gathered, void* stream);

/* perm(A) with EVERY operation in double-double arithmetic (~106 mantissa bits): the engine's
 * counterpart of the reference's extended-precision request
 *   BBFGPermanentCalculator::calculate(mtx, use_extended=true)  sampling/source/BBFGPermanentCalculator.cpp:109-116
 * (x87 long double there).  Same shapes and edge semantics as pqb200_permanent (empty -> 1, rows >=
 * cols+2 -> 0, one row -> product of its entries, rectangular input -> the reference's sum over the
 * `cols` column sums).  About 25x slower than pqb200_permanent; it exists to give ground truth at
 * sizes the CPU reference cannot reach (n=40) and is sharded over `ndev` devices the same way.
 * out_dd = re.hi re.lo im.hi im.lo of the scaled permanent. */
int pqb200_permanent_dd(const double* a, size_t rows, size_t cols, size_t stride, int ndev,
                        double out_dd[4]);

/* The repeated-row permanent with EVERY operation in double-double arithmetic: the counterpart of
 *   BBFGPermanentCalculatorRepeated::calculate(mtx, col_mult, row_mult, use_extended=true / use_inf=true)
 *                                                 sampling/source/BBFGPermanentCalculatorRepeated.cpp:105-141
 * (long double / MPFR there).  Same argument meaning and edge semantics as pqb200_permanent_repeated;
 * out_dd = re.hi re.lo im.hi im.lo.  Slow by design (a straight walk of the reference's n-ary Gray code). */
int pqb200_permanent_repeated_dd(const double* a, size_t rows, size_t cols, size_t stride,
                                 const int64_t* row_mult, const int64_t* col_mult, int ndev,
                                 double out_dd[4]);

/* Many plain permanents in one call: count matrices, matrix k is rows[k] x cols[k] with stride
 * strides[k].  Replaces the list-of-matrices branch of the Python wrapper
 * (sampling/GlynnPermanentCalculator_Wrapper.hpp:204-293). */
int pqb200_permanent_batch(const double* const* a, const size_t* rows, const size_t* cols,
                           const size_t* strides, size_t count, int ndev, double* out /* 2*count */);

/* Repeated-row/column permanent.  row_mult[rows] = output-state occupations, col_mult[cols] =
 * input-state occupations (argument order of the reference, Appendix A.3 of SURVEY.md).  Replaces
 *   BBFGPermanentCalculatorRepeated::calculate(matrix&, col_mult, row_mult, use_extended, use_inf)
 *                                                sampling/source/BBFGPermanentCalculatorRepeated.cpp:74-141
 *   GlynnPermanentCalculatorRepeated<M,P>::calculate(matrix&, input_state, output_state)
 *                                                sampling/source/GlynnPermanentCalculatorRepeated.hpp:49-74
 * Returns PQB200_ERR_MISMATCH when the photon numbers differ (the reference throws std::string,
 * BBFGPermanentCalculatorRepeated.hpp:119-124).
 * Evaluated in the reference's power form (column sums raised to their multiplicity, hpp:263-275):
 * no photon-number cap; the fast kernel takes 45 distinct rows x 48 distinct columns on the two sides after
 * the orientation choice (wider shapes, up to 65 x 128, are served by the double-double walk), and
 * multiplicities <= 254.
 * Precision dispatch: the kernel also returns sum|term|; when the a-posteriori rounding bound
 * eps*(photons+8)*sum|term|/|sum| exceeds 2e-9 the call is re-evaluated in double-double arithmetic
 * (pqb200_permanent_repeated_dd) -- the engine's form of the reference's use_extended / use_inf
 * dispatch (BBFGPermanentCalculatorRepeated.cpp:105-141).  pqb200_last_repeated_bound reports the
 * bound of the calling thread's most recent repeated permanent and whether it was re-evaluated. */
int pqb200_last_repeated_bound(double* bound, int* used_dd);
int pqb200_permanent_repeated(const double* a, size_t rows, size_t cols, size_t stride,
                              const int64_t* row_mult, const int64_t* col_mult, int ndev,
                              double out[2]);

/* `count` repeated permanents of the SAME matrix with different occupation vectors
 * (row_mults: count x rows, col_mults: count x cols, row-major).  Replaces the list-of-states
 * branch of ChinHuhPermanentCalculator_Wrapper.hpp:233-326 and
 * BatchednPermanentCalculator::calculate (sampling/source/BatchedPermanentCalculator.cpp:109-296). */
int pqb200_permanent_repeated_batch(const double* a, size_t rows, size_t cols, size_t stride,
                                    const int64_t* row_mults, const int64_t* col_mults,
                                    size_t count, int ndev, double* out /* 2*count */);

/* One photon step of the Clifford-Clifford B sampler for S shots at once.  For shot s:
 *   pmf[s][m'] ~ | sum_i input[s][i] * perm(U[sample[s] rows, (input[s]-e_i) cols]) * U[m'][i] |^2
 * normalised to sum 1.  Replaces compute_pmf (sampling/simulation_strategies/source/
 * samplingHelperFunctions.cpp:78-342) called once per shot.  Every shot must satisfy
 * sum(input) == sum(sample) + 1 (else PQB200_ERR_MISMATCH, as the reference's calculator throws). */
int pqb200_pmf_layer(const double* u, size_t m, size_t stride, const int64_t* samples /* S x m */,
                     const int64_t* inputs /* S x m */, size_t S, int ndev, double* pmf /* S x m */);

/* The same layer computed as compute_pmf itself does it: one batched kernel job per sub-permanent
 * (k repeated permanents per shot).  pqb200_pmf_layer evaluates all k of them in ONE Gray sweep
 * (Laplace expansion, piquasso-module/src/permanent_laplace.cpp:43-233); this entry point is kept as
 * its cross-check and serves shots with more than 40 photons. */
int pqb200_pmf_layer_batched(const double* u, size_t m, size_t stride, const int64_t* samples,
                             const int64_t* inputs, size_t S, int ndev, double* pmf);

/* Seed the library's process-global std::mt19937.  Replaces seed_random_generator
 * (samplingHelperFunctions.cpp:60-62), which every strategy's seed() forwards to. */
int pqb200_seed(uint64_t seed);

/* fill_r_sample for `shots` shots at once, the photon loop shared by the three B strategies
 * (CGeneralizedCliffordsBSimulationStrategy.cpp:282-316): shot s inserts nphot[s] photons, the k-th
 * one into input mode insert_modes[s*nmax+k] (update_input_by_single_photon with its random choice
 * already resolved), computes the pmf layer and picks the output mode with the uniform draw
 * draws[s*nmax+k] exactly as sample_from_pmf does (samplingHelperFunctions.cpp:350-374).  One kernel
 * thread-block cluster (1-8 CTAs) per shot runs all of its photon steps; the shots are cut into
 * waves that alternate over the devices.  nmax <= 40.
 * samples_out: shots x m. */
int pqb200_fill_r_samples(const double* u, size_t m, size_t stride, size_t shots, size_t nmax,
                          const int32_t* nphot, const uint8_t* insert_modes, const double* draws,
                          int ndev, int64_t* samples_out);

/* The lossless Clifford-Clifford "B" sampler: `shots` samples of the output occupation for the
 * input occupation `input_state` behind interferometer U (m x m).  The global generator is consumed
 * in exactly the reference's order (two draws per photon step, shot after shot), so seeded runs
 * reproduce the reference's shots.  Replaces
 *   CGeneralizedCliffordsBSimulationStrategy::simulate
 *     sampling/simulation_strategies/source/CGeneralizedCliffordsBSimulationStrategy.cpp:137-316
 * samples_out: shots x m, in the reference's C++ generation order.
 * pqb200_simulate_B = pqb200_seed(seed) followed by pqb200_simulate_B_next;
 * pqb200_simulate_B_next continues the generator (a second simulate() on the same strategy). */
int pqb200_simulate_B(const double* u, size_t m, size_t stride, const int64_t* input_state,
                      size_t shots, uint64_t seed, int ndev, int64_t* samples_out);
int pqb200_simulate_B_next(const double* u, size_t m, size_t stride, const int64_t* input_state,
                           size_t shots, int ndev, int64_t* samples_out);

/* Uniform-loss B sampler: per shot the surviving photon number is drawn from
 * Binomial(n, transmittance^2) with one uniform draw, then that many photon steps follow.  Replaces
 *   CGeneralizedCliffordsBUniformLossesSimulationStrategy::simulate
 *     .../CGeneralizedCliffordsBUniformLossesSimulationStrategy.cpp:110-327
 * Continues the global generator (call pqb200_seed first for a seeded run). */
int pqb200_simulate_B_uniform_losses(const double* u, size_t m, size_t stride,
                                     const int64_t* input_state, size_t shots, double transmittance,
                                     int ndev, int64_t* samples_out);

/* Lossy B sampler.  The uniform part of the losses is extracted from the interferometer first
 * (uniform_loss = 1 - sigma_max^2, U <- U / sigma_max: `u` is UPDATED IN PLACE as the reference
 * updates its member matrix, *uniform_loss_out receives the loss); per shot the photons of the
 * first `approximated_modes` modes are put into one random approximated mode and thinned, the
 * other modes are thinned binomially, then the photon loop runs.  shots == 0 only extracts the
 * losses (no device needed).  Replaces
 *   CGeneralizedCliffordsBLossySimulationStrategy::simulate
 *     .../CGeneralizedCliffordsBLossySimulationStrategy.cpp:146-508
 * Continues the global generator. */
int pqb200_simulate_B_lossy(double* u, size_t m, size_t stride, const int64_t* input_state,
                            size_t shots, int approximated_modes, int ndev,
                            double* uniform_loss_out, int64_t* samples_out);

/* Timing of the most recent compute call on this thread (CUDA events around the kernels on device
 * 0 of the call): kernel milliseconds, Gray steps executed, kernels launched. */
typedef struct pqb200_stats {
    double kernel_ms;
    double gray_steps;
    uint64_t launches;
    int ndev;
} pqb200_stats;
int pqb200_last_stats(pqb200_stats* out);

/* Measurement aid (no reference counterpart): sustained issue rate of the FP64 pipe on `device`
 * in thread-instructions per second, from independent chains of op 0 = DFMA, 1 = DADD, 2 = DMUL
 * running for about duration_hint_s seconds.  bench.py uses it as the roofline denominator. */
int pqb200_fp64_peak(int device, int op, double duration_hint_s, double* thread_instr_per_s);

#ifdef __cplusplus
}
#endif
#endif /* PQB200_H */
