// pqb200_capi.cu -- host side of libpqb200.so: the C ABI declared in include/pqb200.h.
//
// Host logic only: argument checks and the reference's edge cases, work planning (chunk length,
// grid, shard ranges), kernel launches on one stream per device, and the double-double combine of
// per-device partials.  There is no CPU compute path: without a CUDA device every call fails.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pqb200.h"
#include "dd.cuh"
#include "launch.h"
#include "plain_host.h"
#include "batch_host.h"
#include "laplace_host.h"
#include "svd_host.h"
#include "repeated_host.h"
#include "rng_replay.h"

using namespace pqb200;

// registered by the glynn_plain_inst.cu units (one per N range, see Makefile PLAIN_RANGES)
#define PQB_DECL_FILL(lo) \
    extern "C" void pqb200_plain_fill_##lo(plain_launch_fn*, plain_occupancy_fn*);
PQB_DECL_FILL(2)
PQB_DECL_FILL(17)
PQB_DECL_FILL(25)
PQB_DECL_FILL(31)
PQB_DECL_FILL(36)
PQB_DECL_FILL(40)
PQB_DECL_FILL(43)
// registered by the glynn_reppow_inst.cu units (one per range of distinct column counts C)
#define PQB_DECL_POW_FILL(lo) extern "C" void pqb200_pow_fill_##lo(pow_launch_fn*);
PQB_DECL_POW_FILL(1)
PQB_DECL_POW_FILL(9)
PQB_DECL_POW_FILL(17)
PQB_DECL_POW_FILL(25)
PQB_DECL_POW_FILL(33)
PQB_DECL_POW_FILL(41)
#define PQB_POW_FILL_CALLS                                                                      \
    pqb200_pow_fill_1(g_pow_launch); pqb200_pow_fill_9(g_pow_launch); pqb200_pow_fill_17(g_pow_launch); \
    pqb200_pow_fill_25(g_pow_launch); pqb200_pow_fill_33(g_pow_launch); pqb200_pow_fill_41(g_pow_launch);

namespace pqb200 {
cudaError_t launch_repeated_dd(const double* d_mat, const double* d_binom, int binom_stride, int D, int N,
                               const int* r, const int* cm, unsigned long long begin, unsigned long long end,
                               int max_blocks, double* partials, unsigned int* counter, double* out,
                               cudaStream_t stream);
cudaError_t launch_plain_dd(const double* d_a, int rows, int n, int L, unsigned long long cb, unsigned long long ce,
                            int grid, double* partials, unsigned int* counter, double* out,
                            cudaStream_t stream);
}

namespace {

constexpr int kMaxPlainN = 45;
constexpr int kMaxChunkLog = 12;  // full re-seed of the running column sums every 2^12 steps
constexpr int kSubChunkLog = 8;   // ... and re-derivation from the per-thread anchor every 2^8 steps
constexpr int kMaxGridBlocks = 4096;
constexpr int kMaxDevices = 64;
constexpr int kSlotsPerDevice = 2;
// FP64 instructions one GPU gets through in about a millisecond: a call with `ndev == 0` (automatic) is
// spread over ceil(work / kWorkPerDevice) devices, so that small calls touch ONE GPU
constexpr double kWorkPerDevice = 2.0e10;

struct LapDeviceBuffers {
    void* nphot = nullptr;   size_t nphot_cap = 0;
    void* ins = nullptr;     size_t ins_cap = 0;
    void* rnd = nullptr;     size_t rnd_cap = 0;
    void* samples = nullptr; size_t samples_cap = 0;
    void* sin = nullptr;     size_t sin_cap = 0;
    void* cin = nullptr;     size_t cin_cap = 0;
    void* pmf = nullptr;     size_t pmf_cap = 0;
    void* order = nullptr;   size_t order_cap = 0;
    void* h_stage = nullptr; size_t h_stage_cap = 0;  // pinned host staging of the sampler's waves
    unsigned long long* terms = nullptr;  // [0] Gray terms swept, [1] work-queue head
    bool ready = false;
};

// One WORK SLOT of one device: a stream plus every workspace a call needs.  A call owns the slots it
// runs on (slot mutex held from the first enqueue to the last synchronise); calls on different
// devices -- and up to kSlotsPerDevice calls on the same device -- run concurrently.
struct DeviceCtx {
    int dev = -1;
    int slot = 0;
    int sm_count = 0;
    std::mutex mu;
    cudaStream_t stream = nullptr;
    double* partials = nullptr;  // kMaxGridBlocks x 4
    unsigned int* counter = nullptr;
    double* d_out = nullptr;  // 8 doubles
    double* h_out = nullptr;  // pinned, 8 doubles
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    std::vector<cudaEvent_t> evs;  // extra timing events (calls that time several separate launches)
    // Recorded behind work that an asynchronous entry point (pqb200_permanent_shard with d_out,
    // pqb200_permanent_nccl_async) left in flight on a CALLER's stream: every later use of the slot's
    // workspaces is ordered behind it with cudaStreamWaitEvent.
    cudaEvent_t busy = nullptr;
    bool busy_pending = false;
    bool dirty = false;  // a call failed half-way: drain the stream and reset the counter before reuse
    int occ[kPlainTableSize];
    // repeated-row kernel tables (device): base vectors, binomials
    void* rep_base = nullptr;    size_t rep_base_cap = 0;
    void* rep_binom = nullptr;   size_t rep_binom_cap = 0;
    // double-double entry points: matrix and binomial table
    void* dd_a = nullptr;        size_t dd_a_cap = 0;
    void* dd_b = nullptr;        size_t dd_b_cap = 0;
    // batch path: growable device buffers (matrix data, job descriptors, results, pmf scratch)
    void* b_u = nullptr;      size_t b_u_cap = 0;
    void* b_jobs = nullptr;   size_t b_jobs_cap = 0;
    void* b_out = nullptr;    size_t b_out_cap = 0;
    void* b_weight = nullptr; size_t b_weight_cap = 0;
    void* b_pmf = nullptr;    size_t b_pmf_cap = 0;
    bool batch_ready = false;
    LapDeviceBuffers lap;
};

// g_mu guards library-wide state only (device discovery and slot creation, the default device
// count, the NCCL communicator, the reference's global random generator); it is never held while a
// kernel runs.
std::mutex g_mu;
DeviceCtx g_slots[kMaxDevices][kSlotsPerDevice];
bool g_tables_ready = false;
plain_launch_fn g_plain_launch[kPlainTableSize];
plain_occupancy_fn g_plain_occ[kPlainTableSize];
pow_launch_fn g_pow_launch[kPowTableSize];

thread_local std::string g_err;
thread_local pqb200_stats g_stats = {0.0, 0.0, 0, 0};

// PQB200_TRACE=1: host-side phase timings of the batch paths on stderr (tuning aid)
bool trace_on() {
    static int on = -1;
    if (on < 0) on = getenv("PQB200_TRACE") ? 1 : 0;
    return on == 1;
}
double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                        \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return fail(PQB200_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(e__) + \
                                             " in " #expr);                                   \
    } while (0)

void fill_tables() {
    if (g_tables_ready) return;
    memset(g_plain_launch, 0, sizeof(g_plain_launch));
    memset(g_plain_occ, 0, sizeof(g_plain_occ));
    pqb200_plain_fill_2(g_plain_launch, g_plain_occ);
    pqb200_plain_fill_17(g_plain_launch, g_plain_occ);
    pqb200_plain_fill_25(g_plain_launch, g_plain_occ);
    pqb200_plain_fill_31(g_plain_launch, g_plain_occ);
    pqb200_plain_fill_36(g_plain_launch, g_plain_occ);
    pqb200_plain_fill_40(g_plain_launch, g_plain_occ);
    pqb200_plain_fill_43(g_plain_launch, g_plain_occ);
    memset(g_pow_launch, 0, sizeof(g_pow_launch));
    PQB_POW_FILL_CALLS
    g_tables_ready = true;
}

int g_visible = -1;      // CUDA devices visible to this process
int g_default_ndev = 0;  // devices one call may be sharded over when the caller passes ndev == 0

int grow(void** p, size_t* cap, size_t need) {
    if (need <= *cap) return PQB200_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    const size_t want = std::max(need, (size_t)1 << 16) * 3 / 2;
    CUDA_TRY(cudaMalloc(p, want));
    *cap = want;
    return PQB200_OK;
}


// caller holds g_mu.  Counts the visible devices (no context is created).
int ensure_visible() {
    fill_tables();
    if (g_visible < 0) {
        int visible = 0;
        cudaError_t e = cudaGetDeviceCount(&visible);
        if (e != cudaSuccess || visible <= 0)
            return fail(PQB200_ERR_CUDA,
                        std::string("no usable CUDA device (libpqb200 has no CPU fallback): ") +
                            (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
        g_visible = std::min(visible, kMaxDevices);
    }
    return PQB200_OK;
}

// caller holds g_mu.  Creates the slot (stream, workspaces) on first use only, so that a
// one-process-per-GPU job touches nothing but its own device.  Changes the calling thread's current
// device (callers restore it: Lease).
int get_device(int dev, int slot, DeviceCtx** out) {
    int rc = ensure_visible();
    if (rc) return rc;
    if (dev < 0 || dev >= g_visible) return fail(PQB200_ERR_ARG, "device index out of range");
    if (slot < 0 || slot >= kSlotsPerDevice) return fail(PQB200_ERR_ARG, "slot index out of range");
    DeviceCtx& c = g_slots[dev][slot];
    if (c.dev < 0) {
        CUDA_TRY(cudaSetDevice(dev));
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
        if (prop.major < 10)
            return fail(PQB200_ERR_CUDA, std::string("device ") + prop.name +
                                             " is not sm_100 class; this library is built for B200 only");
        c.sm_count = prop.multiProcessorCount;
        c.slot = slot;
        CUDA_TRY(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaMalloc(&c.partials, sizeof(double) * 8 * kMaxGridBlocks));
        CUDA_TRY(cudaMalloc(&c.counter, sizeof(unsigned int)));
        CUDA_TRY(cudaMemset(c.counter, 0, sizeof(unsigned int)));
        CUDA_TRY(cudaMalloc(&c.d_out, sizeof(double) * 8));
        CUDA_TRY(cudaMallocHost(&c.h_out, sizeof(double) * 8));
        CUDA_TRY(cudaEventCreate(&c.e0));
        CUDA_TRY(cudaEventCreate(&c.e1));
        CUDA_TRY(cudaEventCreateWithFlags(&c.busy, cudaEventDisableTiming));
        for (int i = 0; i < kPlainTableSize; i++) c.occ[i] = 0;
        c.dev = dev;
    }
    if (out) *out = &c;
    return PQB200_OK;
}

// caller holds g_mu
int init_locked(int ndev) {
    int rc = ensure_visible();
    if (rc) return rc;
    if (ndev <= 0 || ndev > g_visible) ndev = g_visible;
    g_default_ndev = ndev;
    return PQB200_OK;
}

// The slots one call runs on, locked; released (and the caller's current CUDA device restored) when
// the lease goes out of scope.  A call that returns early with an error leaves its slots marked
// dirty: the next owner drains them before use.
struct Lease {
    std::vector<DeviceCtx*> slots;
    int saved_dev = -1;
    bool ok = false;
    Lease() = default;
    Lease(const Lease&) = delete;
    Lease& operator=(const Lease&) = delete;
    ~Lease() { release(); }
    void release() {
        for (DeviceCtx* c : slots) {
            if (!ok) c->dirty = true;
            c->mu.unlock();
        }
        slots.clear();
        int cur = -1;
        if (saved_dev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != saved_dev) cudaSetDevice(saved_dev);
        saved_dev = -1;
    }
    int size() const { return (int)slots.size(); }
    DeviceCtx& operator[](int i) const { return *slots[(size_t)i]; }
    int done(int rc) {
        ok = (rc == PQB200_OK);
        return rc;
    }
};

// after the slot mutex has been taken
int slot_prepare(DeviceCtx& c) {
    if (c.dirty || c.busy_pending) CUDA_TRY(cudaSetDevice(c.dev));
    if (c.dirty) {
        cudaStreamSynchronize(c.stream);
        cudaGetLastError();
        CUDA_TRY(cudaMemsetAsync(c.counter, 0, sizeof(unsigned int), c.stream));
        c.dirty = false;
    }
    if (c.busy_pending) {
        if (cudaEventQuery(c.busy) == cudaSuccess) c.busy_pending = false;
        else CUDA_TRY(cudaStreamWaitEvent(c.stream, c.busy, 0));
    }
    return PQB200_OK;
}

// Resolve the `ndev` argument of a compute call and lock the slots it runs on.
//   ndev >= 2 : devices 0..ndev-1 (slot 0 of each, taken in ascending order);
//   ndev == 1 : the calling thread's current CUDA device (cudaSetDevice / torch.cuda.set_device; 0 if
//               never set), whichever of its slots is free;
//   ndev == 0 : automatic -- ceil(work / kWorkPerDevice) devices, at most pqb200_init's count; a call
//               that needs ONE device prefers the current device and moves to another initialised
//               device when both of its slots are busy (concurrent callers spread over the GPUs).
// `work` = estimated FP64 instructions of the call.
int acquire(int ndev, double work, Lease& ls) {
    int used, pref = 0;
    bool roam = false;
    cudaGetDevice(&ls.saved_dev);  // before anything below changes it: restored when the lease ends
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (g_default_ndev == 0) {
            int rc = init_locked(0);
            if (rc) return rc;
        }
        if (ndev <= 0) {
            const double want = std::ceil(work / kWorkPerDevice);
            used = want >= (double)g_default_ndev ? g_default_ndev : std::max(1, (int)want);
            roam = true;
        } else {
            used = std::min(ndev, g_visible);
        }
        if (used == 1) {
            pref = ls.saved_dev;
            if (pref < 0 || pref >= (roam ? g_default_ndev : g_visible)) pref = 0;
            for (int s = 0; s < kSlotsPerDevice; s++) {
                int rc = get_device(pref, s, nullptr);
                if (rc) return rc;
            }
        } else {
            for (int d = 0; d < used; d++) {
                int rc = get_device(d, 0, nullptr);
                if (rc) return rc;
            }
        }
    }
    if (used == 1) {
        DeviceCtx* got = nullptr;
        for (int s = 0; s < kSlotsPerDevice && !got; s++)
            if (g_slots[pref][s].mu.try_lock()) got = &g_slots[pref][s];
        if (!got && roam) {
            for (int k = 1; k < g_default_ndev && !got; k++) {
                const int d = (pref + k) % g_default_ndev;
                for (int s = 0; s < kSlotsPerDevice && !got; s++) {
                    DeviceCtx& c = g_slots[d][s];
                    if (c.dev < 0) {  // not created yet: create it (first use of that GPU by this process)
                        std::lock_guard<std::mutex> lk(g_mu);
                        if (get_device(d, s, nullptr)) continue;
                    }
                    if (c.mu.try_lock()) got = &c;
                }
            }
        }
        if (!got) {
            g_slots[pref][0].mu.lock();
            got = &g_slots[pref][0];
        }
        ls.slots.push_back(got);
    } else {
        for (int d = 0; d < used; d++) {
            g_slots[d][0].mu.lock();
            ls.slots.push_back(&g_slots[d][0]);
        }
    }
    for (DeviceCtx* c : ls.slots) {
        int rc = slot_prepare(*c);
        if (rc) return rc;
    }
    return PQB200_OK;
}

// lock one given slot (the one-process-per-GPU entry points name their device)
int acquire_device(int device, Lease& ls) {
    cudaGetDevice(&ls.saved_dev);
    {
        std::lock_guard<std::mutex> lk(g_mu);
        int rc = get_device(device, 0, nullptr);
        if (rc) return rc;
    }
    g_slots[device][0].mu.lock();
    ls.slots.push_back(&g_slots[device][0]);
    return slot_prepare(*ls.slots[0]);
}

// Order a launch on a CALLER's stream behind whatever an earlier asynchronous call left in flight
// on this slot's workspaces.
int order_behind_async(DeviceCtx& c, cudaStream_t st) {
    if (c.busy_pending && st != c.stream) CUDA_TRY(cudaStreamWaitEvent(st, c.busy, 0));
    return PQB200_OK;
}
int mark_async(DeviceCtx& c, cudaStream_t st) {
    CUDA_TRY(cudaEventRecord(c.busy, st));
    c.busy_pending = true;
    return PQB200_OK;
}

struct PlainPlan {
    int N;   // kernel template parameter (columns incl. padding)
    int L;   // log2 chunk length
    int K;   // log2 sub-chunk length (anchor re-seed period)
    unsigned long long nchunks;
};

// Choose the chunk length: minimise  waves * (seed cost + 2^L * step cost)  for the busiest device.
PlainPlan plan_plain(int rows, int cols, int max_resident_threads, unsigned long long shards) {
    PlainPlan pl;
    pl.N = plain_kernel_cols(rows, cols);
    const int lmax = std::min(rows - 1, kMaxChunkLog);
    double best = 1e300;
    pl.L = lmax;
    // L == 1 only for rows == 2 (a single chunk): the kernel derives every toggle direction from the
    // local step, which needs L >= 2 as soon as there is more than one chunk
    const int lmin = std::min(2, lmax);
    for (int L = lmin; L <= lmax; L++) {
        const unsigned long long C = 1ull << (rows - 1 - L);
        const unsigned long long per_shard = (C + shards - 1) / shards;
        const unsigned long long threads =
            std::min<unsigned long long>((unsigned long long)max_resident_threads,
                                         ((per_shard + kPlainThreads - 1) / kPlainThreads) * kPlainThreads);
        const unsigned long long waves = (per_shard + threads - 1) / threads;
        const double cost = (double)waves * (2.0 * rows * pl.N + (double)(1ull << L) * (6.0 * pl.N + 5.0));
        if (cost < best * 0.999) {  // prefer the shorter chunk on ties (less rounding drift)
            best = cost;
            pl.L = L;
        }
    }
    if (const char* f = getenv("PQB200_FORCE_L")) {  // tuning / debugging aid
        const int L = atoi(f);
        if (L >= lmin && L <= rows - 1) pl.L = L;
    }
    pl.nchunks = 1ull << (rows - 1 - pl.L);
    pl.K = std::min(pl.L, kSubChunkLog);
    if (const char* f = getenv("PQB200_FORCE_K")) {  // tuning / debugging aid
        const int K = atoi(f);
        if (K >= std::min(2, pl.L) && K <= pl.L) pl.K = K;
    }
    return pl;
}

int plain_occupancy(DeviceCtx& c, int N) {
    if (c.occ[N] == 0) {
        cudaSetDevice(c.dev);
        c.occ[N] = g_plain_occ[N]();
    }
    return c.occ[N];
}

// Launch the chunk range [cb, ce) of the plain kernel on slot c (async on `stream`), result to `out`.
int launch_plain_range(DeviceCtx& c, const double* a, size_t rows, size_t cols, size_t stride,
                       const PlainPlan& pl, unsigned long long cb, unsigned long long ce,
                       double* out, cudaStream_t stream) {
    CUDA_TRY(cudaSetDevice(c.dev));
    int rc = order_behind_async(c, stream);
    if (rc) return rc;
    if (ce <= cb) {
        CUDA_TRY(cudaMemsetAsync(out, 0, 4 * sizeof(double), stream));
        return PQB200_OK;
    }
    const int occ = plain_occupancy(c, pl.N);
    if (occ <= 0) return fail(PQB200_ERR_CUDA, "occupancy query failed for the plain kernel");
    PlainLaunchArgs x;
    x.a = a;
    x.rows = rows;
    x.cols = cols;
    x.stride = stride;
    x.L = pl.L;
    x.K = pl.K;
    x.chunk_begin = cb;
    x.chunk_end = ce;
    const unsigned long long want = (ce - cb + kPlainThreads - 1) / kPlainThreads;
    x.grid = (int)std::min<unsigned long long>(want, (unsigned long long)std::min(kMaxGridBlocks, occ * c.sm_count));
    x.partials = c.partials;
    x.counter = c.counter;
    x.out = out;
    x.stream = stream;
    CUDA_TRY(g_plain_launch[pl.N](x));
    return PQB200_OK;
}

int check_plain_args(const double* a, size_t rows, size_t cols, size_t stride) {
    if (!a) return fail(PQB200_ERR_ARG, "matrix pointer is null");
    if (stride < cols) return fail(PQB200_ERR_ARG, "stride is smaller than the number of columns");
    if (rows >= cols + 2) return PQB200_OK;  // the reference returns 0 for these whatever the size
    if (std::max(rows, cols) > (size_t)kMaxPlainN ||
        (rows >= 2 && plain_kernel_cols((int)rows, (int)cols) > kMaxPlainN))
        return fail(PQB200_ERR_ARG, "matrix dimension exceeds 45 (kernel-parameter resident matrix limit of this build)");
    return PQB200_OK;
}

// reference edge cases (BBFGPermanentCalculator.hpp:90-95); returns true when `out` is final
bool plain_edge_case(const double* a, size_t rows, size_t cols, size_t stride, double out[2]) {
    if (rows == 0 || cols == 0) {
        out[0] = 1.0;
        out[1] = 0.0;
        return true;
    }
    if (rows >= cols + 2) {
        out[0] = 0.0;
        out[1] = 0.0;
        return true;
    }
    if (rows == 1) {
        // a single row has one delta vector: the "sum" is the product of the row's entries
        double re = 1.0, im = 0.0;
        for (size_t j = 0; j < cols; j++) {
            const double xr = a[2 * j], xi = a[2 * j + 1];
            const double r = re * xr - im * xi, i = re * xi + im * xr;
            re = r;
            im = i;
        }
        (void)stride;
        out[0] = re;
        out[1] = im;
        return true;
    }
    return false;
}

void combine_partials(const double* parts, size_t count, size_t rows, double out[2]) {
    cdd tot = cdd_zero();
    for (size_t k = 0; k < count; k++) {
        cdd x;
        x.re.hi = parts[4 * k + 0];
        x.re.lo = parts[4 * k + 1];
        x.im.hi = parts[4 * k + 2];
        x.im.lo = parts[4 * k + 3];
        cdd_add(tot, x);
    }
    const int e = -(int)(rows - 1);
    out[0] = ldexp(tot.re.hi, e) + ldexp(tot.re.lo, e);
    out[1] = ldexp(tot.im.hi, e) + ldexp(tot.im.lo, e);
}

// FP64 instructions of one plain permanent (5.5N - 6.5 per Gray step)
double plain_work(size_t rows, size_t cols) {
    if (rows < 2 || rows >= cols + 2) return 0.0;
    return ldexp(1.0, (int)rows - 1) * 5.5 * (double)plain_kernel_cols((int)rows, (int)cols);
}

// FP64 instructions of S photon steps with at most `photons` photons each: the bound for all-distinct
// occupations (2^photons x 14 x photons) scaled by the measured effect of collisions (C5: 1.1e8 Gray terms
// per 1000 shots instead of 5e8); used only to choose the device count of an automatic call
double laplace_work(double S, double photons) {
    return 0.2 * S * ldexp(14.0 * photons, (int)std::min(photons, 60.0) - 1);
}

const RepLimits kRepLimits = {kPowMaxDigits, kPowMaxC, 254, kPowBufBytes, 32, kMaxPlainN};
// the double-double walk reads the matrix from global memory and keeps its sums in local memory
const RepLimits kRepDDLimits = {kDDRepDigits, kDDRepCols, 254, 1 << 30, 0, kDDRepCols};

}  // namespace

extern "C" {

static void lap_release(DeviceCtx& c);
static void nccl_release_locked();

const char* pqb200_last_error(void) { return g_err.c_str(); }
const char* pqb200_version(void) { return "pqb200 0.2 (sm_100a)"; }

int pqb200_init(int ndev) {
    std::lock_guard<std::mutex> lk(g_mu);
    return init_locked(ndev);
}

void pqb200_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    int saved = -1;
    cudaGetDevice(&saved);
    nccl_release_locked();
    for (int d = 0; d < kMaxDevices; d++)
        for (int sl = 0; sl < kSlotsPerDevice; sl++) {
            DeviceCtx& c = g_slots[d][sl];
            if (c.dev < 0) continue;
            std::lock_guard<std::mutex> sk(c.mu);
            cudaSetDevice(c.dev);
            cudaStreamSynchronize(c.stream);
            cudaFree(c.partials);
            cudaFree(c.counter);
            cudaFree(c.d_out);
            cudaFree(c.rep_base);
            cudaFree(c.rep_binom);
            cudaFree(c.dd_a);
            cudaFree(c.dd_b);
            cudaFree(c.b_u);
            cudaFree(c.b_jobs);
            cudaFree(c.b_out);
            cudaFree(c.b_weight);
            cudaFree(c.b_pmf);
            lap_release(c);
            cudaFreeHost(c.h_out);
            cudaEventDestroy(c.e0);
            cudaEventDestroy(c.e1);
            cudaEventDestroy(c.busy);
            for (cudaEvent_t e : c.evs) cudaEventDestroy(e);
            c.evs.clear();
            cudaStreamDestroy(c.stream);
            c.partials = nullptr; c.counter = nullptr; c.d_out = nullptr; c.h_out = nullptr;
            c.rep_base = c.rep_binom = c.dd_a = c.dd_b = nullptr;
            c.rep_base_cap = c.rep_binom_cap = c.dd_a_cap = c.dd_b_cap = 0;
            c.b_u = c.b_jobs = c.b_out = c.b_weight = c.b_pmf = nullptr;
            c.b_u_cap = c.b_jobs_cap = c.b_out_cap = c.b_weight_cap = c.b_pmf_cap = 0;
            c.batch_ready = false;
            c.busy_pending = c.dirty = false;
            c.stream = nullptr;
            c.dev = -1;
        }
    g_visible = -1;
    g_default_ndev = 0;
    if (saved >= 0) cudaSetDevice(saved);
}

int pqb200_device_count(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_default_ndev == 0) {
        if (init_locked(0)) return 0;
    }
    return g_default_ndev;
}

int pqb200_last_stats(pqb200_stats* out) {
    if (!out) return fail(PQB200_ERR_ARG, "stats pointer is null");
    *out = g_stats;
    return PQB200_OK;
}

// perm of a rows x cols matrix (2 <= rows <= cols+1) over the leased devices
static int plain_multi_device(Lease& ls, const double* a, size_t rows, size_t cols, size_t stride,
                              size_t scale_rows, double out[2]) {
    int rc;
    int used = ls.size();
    const int max_threads = ls[0].sm_count * kPlainThreads;
    PlainPlan pl = plan_plain((int)rows, (int)cols, max_threads, (unsigned long long)used);
    if ((unsigned long long)used > pl.nchunks) used = (int)pl.nchunks;
    std::vector<double> parts(4 * (size_t)used);
    CUDA_TRY(cudaSetDevice(ls[0].dev));
    CUDA_TRY(cudaEventRecord(ls[0].e0, ls[0].stream));
    for (int d = 0; d < used; d++) {
        DeviceCtx& c = ls[d];
        const unsigned long long cb = pl.nchunks * (unsigned long long)d / (unsigned long long)used;
        const unsigned long long ce = pl.nchunks * (unsigned long long)(d + 1) / (unsigned long long)used;
        rc = launch_plain_range(c, a, rows, cols, stride, pl, cb, ce, c.d_out, c.stream);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(c.h_out, c.d_out, 4 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        if (d == 0) CUDA_TRY(cudaEventRecord(c.e1, c.stream));
    }
    for (int d = 0; d < used; d++) {
        DeviceCtx& c = ls[d];
        CUDA_TRY(cudaSetDevice(c.dev));
        CUDA_TRY(cudaStreamSynchronize(c.stream));
        memcpy(&parts[4 * (size_t)d], c.h_out, 4 * sizeof(double));
    }
    float ms = 0.f;
    CUDA_TRY(cudaSetDevice(ls[0].dev));
    CUDA_TRY(cudaEventElapsedTime(&ms, ls[0].e0, ls[0].e1));
    g_stats.kernel_ms += ms;
    g_stats.gray_steps += ldexp(1.0, (int)rows - 1);
    g_stats.launches += (uint64_t)used;
    g_stats.ndev = used;
    combine_partials(parts.data(), (size_t)used, scale_rows, out);
    return PQB200_OK;
}

int pqb200_permanent(const double* a, size_t rows, size_t cols, size_t stride, int ndev,
                     double out[2]) {
    if (!out) return fail(PQB200_ERR_ARG, "output pointer is null");
    if (rows != 0 && cols != 0) {
        int rc = check_plain_args(a, rows, cols, stride);
        if (rc) return rc;
    }
    Lease ls;
    // the library must be usable (CUDA present) even for the trivial cases: no silent CPU path
    int rc = acquire(ndev, plain_work(rows, cols), ls);
    if (rc) return rc;
    g_stats = {0.0, 0.0, 0, ls.size()};
    if (plain_edge_case(a, rows, cols, stride, out)) return ls.done(PQB200_OK);
    return ls.done(plain_multi_device(ls, a, rows, cols, stride, rows, out));
}

int pqb200_permanent_dd(const double* a, size_t rows, size_t cols, size_t stride, int ndev,
                        double out_dd[4]) {
    if (!out_dd) return fail(PQB200_ERR_ARG, "output pointer is null");
    if (rows != 0 && cols != 0) {
        int rc = check_plain_args(a, rows, cols, stride);
        if (rc) return rc;
    }
    Lease ls;
    int rc = acquire(ndev, 40.0 * plain_work(rows, cols), ls);
    if (rc) return rc;
    int used = ls.size();
    g_stats = {0.0, 0.0, 0, used};
    {   // the reference's trivial branches (BBFGPermanentCalculator.hpp:90-95) come before everything else
        double e[2];
        if (plain_edge_case(a, rows, cols, stride, e)) {
            out_dd[0] = e[0];
            out_dd[1] = 0.0;
            out_dd[2] = e[1];
            out_dd[3] = 0.0;
            return ls.done(PQB200_OK);
        }
    }
    // rows != cols: the reference's sum runs over the `cols` column sums; as in the fast kernel the
    // missing columns are the constant e_0 (column sum == 1 for every delta vector)
    const int R = (int)rows, n = plain_kernel_cols((int)rows, (int)cols);
    const int L = std::min(R - 1, kMaxChunkLog);
    const unsigned long long nchunks = 1ull << (R - 1 - L);
    if ((unsigned long long)used > nchunks) used = (int)nchunks;
    std::vector<double> dense(2 * (size_t)R * n, 0.0);
    for (int i = 0; i < R; i++) memcpy(&dense[2 * (size_t)i * n], a + 2 * i * stride, 16 * cols);
    for (int j = (int)cols; j < n; j++) dense[2 * (size_t)j] = 1.0;
    std::vector<double> parts(4 * (size_t)used);
    g_stats = {0.0, ldexp(1.0, R - 1), (uint64_t)used, used};
    CUDA_TRY(cudaSetDevice(ls[0].dev));
    CUDA_TRY(cudaEventRecord(ls[0].e0, ls[0].stream));
    for (int d = 0; d < used; d++) {
        DeviceCtx& c = ls[d];
        CUDA_TRY(cudaSetDevice(c.dev));
        rc = grow(&c.dd_a, &c.dd_a_cap, 16 * (size_t)R * n);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(c.dd_a, dense.data(), 16 * (size_t)R * n, cudaMemcpyHostToDevice, c.stream));
        const unsigned long long cb = nchunks * (unsigned long long)d / (unsigned long long)used;
        const unsigned long long ce = nchunks * (unsigned long long)(d + 1) / (unsigned long long)used;
        const unsigned long long want = (ce - cb + kPlainThreads - 1) / kPlainThreads;
        const int grid = (int)std::min<unsigned long long>(want, (unsigned long long)c.sm_count);
        CUDA_TRY(launch_plain_dd((const double*)c.dd_a, R, n, L, cb, ce, grid, c.partials, c.counter, c.d_out, c.stream));
        CUDA_TRY(cudaMemcpyAsync(c.h_out, c.d_out, 4 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        if (d == 0) CUDA_TRY(cudaEventRecord(c.e1, c.stream));
    }
    for (int d = 0; d < used; d++) {
        DeviceCtx& c = ls[d];
        CUDA_TRY(cudaSetDevice(c.dev));
        CUDA_TRY(cudaStreamSynchronize(c.stream));  // `dense` is a host temporary
        memcpy(&parts[4 * (size_t)d], c.h_out, 4 * sizeof(double));
    }
    float ms = 0.f;
    CUDA_TRY(cudaSetDevice(ls[0].dev));
    CUDA_TRY(cudaEventElapsedTime(&ms, ls[0].e0, ls[0].e1));
    g_stats.kernel_ms = ms;
    cdd tot = cdd_zero();
    for (int d = 0; d < used; d++) {
        cdd x;
        x.re.hi = parts[4 * d + 0];
        x.re.lo = parts[4 * d + 1];
        x.im.hi = parts[4 * d + 2];
        x.im.lo = parts[4 * d + 3];
        cdd_add(tot, x);
    }
    const int e = -(R - 1);  // exact power-of-two scaling of both parts
    out_dd[0] = ldexp(tot.re.hi, e);
    out_dd[1] = ldexp(tot.re.lo, e);
    out_dd[2] = ldexp(tot.im.hi, e);
    out_dd[3] = ldexp(tot.im.lo, e);
    return ls.done(PQB200_OK);
}

static double rep_total_steps(const RepPrepared& pr) {
    double t = 1.0;
    for (int m : pr.r) t *= (double)(m + 1);
    return t;
}

// the double-double walk of a prepared repeated permanent over the leased devices
static int repeated_dd_prepared(Lease& ls, const RepPrepared& pr, double out_dd[4]) {
    int used = ls.size();
    const int D = pr.R - 1, C = pr.C;
    const double total_d = rep_total_steps(pr);
    if (total_d > 1.8e19) return fail(PQB200_ERR_ARG, "Gray space exceeds 2^64 steps");
    unsigned long long total = 1;
    int maxr = 1;
    for (int d = 0; d < D; d++) {
        total *= (unsigned long long)(pr.r[d] + 1);
        maxr = std::max(maxr, pr.r[d]);
    }
    const int bstride = maxr + 1;
    std::vector<double> binom((size_t)std::max(D, 1) * bstride * 2, 0.0);
    for (int d = 0; d < D; d++)
        for (int k = 0; k <= pr.r[d]; k++)
            binom_dd(pr.r[d], k, binom[2 * ((size_t)d * bstride + k)], binom[2 * ((size_t)d * bstride + k) + 1]);
    if ((unsigned long long)used > total) used = (int)total;
    g_stats.gray_steps += total_d;
    g_stats.launches += (uint64_t)used;
    CUDA_TRY(cudaSetDevice(ls[0].dev));
    CUDA_TRY(cudaEventRecord(ls[0].e0, ls[0].stream));
    for (int d = 0; d < used; d++) {
        DeviceCtx& c = ls[d];
        CUDA_TRY(cudaSetDevice(c.dev));
        int rc = grow(&c.dd_a, &c.dd_a_cap, sizeof(double) * pr.mat.size());
        if (rc) return rc;
        if ((rc = grow(&c.dd_b, &c.dd_b_cap, sizeof(double) * binom.size()))) return rc;
        CUDA_TRY(cudaMemcpyAsync(c.dd_a, pr.mat.data(), sizeof(double) * pr.mat.size(), cudaMemcpyHostToDevice, c.stream));
        CUDA_TRY(cudaMemcpyAsync(c.dd_b, binom.data(), sizeof(double) * binom.size(), cudaMemcpyHostToDevice, c.stream));
        const unsigned long long b = (unsigned long long)((__uint128_t)total * (unsigned)d / (unsigned)used);
        const unsigned long long e = (unsigned long long)((__uint128_t)total * (unsigned)(d + 1) / (unsigned)used);
        CUDA_TRY(launch_repeated_dd((const double*)c.dd_a, (const double*)c.dd_b, bstride, D, C, pr.r.data(),
                                    pr.cm.data(), b, e, c.sm_count, c.partials, c.counter, c.d_out, c.stream));
        CUDA_TRY(cudaMemcpyAsync(c.h_out, c.d_out, 4 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        if (d == 0) CUDA_TRY(cudaEventRecord(c.e1, c.stream));
    }
    cdd tot = cdd_zero();
    for (int d = 0; d < used; d++) {
        DeviceCtx& c = ls[d];
        CUDA_TRY(cudaSetDevice(c.dev));
        CUDA_TRY(cudaStreamSynchronize(c.stream));
        cdd x;
        x.re.hi = c.h_out[0];
        x.re.lo = c.h_out[1];
        x.im.hi = c.h_out[2];
        x.im.lo = c.h_out[3];
        cdd_add(tot, x);
    }
    float ms = 0.f;
    CUDA_TRY(cudaSetDevice(ls[0].dev));
    CUDA_TRY(cudaEventElapsedTime(&ms, ls[0].e0, ls[0].e1));
    g_stats.kernel_ms += ms;
    const int e = -(pr.photons - 1);
    out_dd[0] = ldexp(tot.re.hi, e);
    out_dd[1] = ldexp(tot.re.lo, e);
    out_dd[2] = ldexp(tot.im.hi, e);
    out_dd[3] = ldexp(tot.im.lo, e);
    return PQB200_OK;
}

static int check_repeated_args(const double* a, size_t rows, size_t cols, size_t stride,
                               const int64_t* row_mult, const int64_t* col_mult) {
    if ((rows && !row_mult) || (cols && !col_mult)) return fail(PQB200_ERR_ARG, "multiplicity pointer is null");
    if (rows && cols && !a) return fail(PQB200_ERR_ARG, "matrix pointer is null");
    if (rows && cols && stride < cols) return fail(PQB200_ERR_ARG, "stride is smaller than the number of columns");
    return PQB200_OK;
}

int pqb200_permanent_repeated_dd(const double* a, size_t rows, size_t cols, size_t stride,
                                 const int64_t* row_mult, const int64_t* col_mult, int ndev,
                                 double out_dd[4]) {
    if (!out_dd) return fail(PQB200_ERR_ARG, "output pointer is null");
    int rc = check_repeated_args(a, rows, cols, stride, row_mult, col_mult);
    if (rc) return rc;
    RepPrepared pr;
    std::string err;
    rc = rep_prepare(a, rows, cols, stride, row_mult, col_mult, kRepDDLimits, pr, err);
    if (rc) return fail(rc == 4 ? PQB200_ERR_ARG : rc, err);
    Lease ls;
    rc = acquire(ndev, pr.trivial ? 0.0 : rep_total_steps(pr) * 200.0 * pr.photons, ls);
    if (rc) return rc;
    g_stats = {0.0, 0.0, 0, ls.size()};
    out_dd[1] = out_dd[3] = 0.0;
    if (pr.trivial) {
        out_dd[0] = pr.value[0];
        out_dd[2] = pr.value[1];
        return ls.done(PQB200_OK);
    }
    return ls.done(repeated_dd_prepared(ls, pr, out_dd));
}

// A-posteriori rounding-error bound of the fast repeated kernel: eps * (photons + 8) * sum|term| /
// |sum|.  Every term is a product of `photons` FP64 factors, each carrying a few roundings of its
// column sum; the bound treats them as adding up linearly (worst case), so the true error is well
// below it: measured bound / error = 35 ... 800 (profiles/repeated_bound_calibration_r2.json), so a
// threshold of 2e-9 keeps the fast kernel's error below 6e-11.  Above kRepeatedDDThreshold the call is
// re-evaluated in double-double arithmetic -- what the reference does up front by size / request
// (BBFGPermanentCalculatorRepeated.cpp:105-141 use_extended / use_inf) is decided here by measurement.
constexpr double kRepeatedDDThreshold = 2.0e-9;
thread_local double g_last_bound = 0.0;
thread_local int g_last_dd = 0;

// one repeated-row permanent over the leased devices
static int repeated_locked(Lease& ls, const double* a, size_t rows, size_t cols, size_t stride,
                           const int64_t* row_mult, const int64_t* col_mult, double out[2]) {
    int used = ls.size();
    RepPrepared pr;
    std::string err;
    int rc = rep_prepare(a, rows, cols, stride, row_mult, col_mult, kRepLimits, pr, err);
    g_last_bound = 0.0;
    g_last_dd = 0;
    if (rc == 4) {
        // beyond the fast kernel's shape limits (more than 48 distinct columns on the product side, more
        // than 45 rows on the Gray side): the double-double walk takes up to 65 x 128, slowly but exactly
        RepPrepared wide;
        std::string err2;
        if (rep_prepare(a, rows, cols, stride, row_mult, col_mult, kRepDDLimits, wide, err2) != 0)
            return fail(PQB200_ERR_ARG, err);
        if (rep_total_steps(wide) > 1e13)
            return fail(PQB200_ERR_ARG, err + "; the double-double fallback would need more than 1e13 Gray steps");
        double dd4[4];
        rc = repeated_dd_prepared(ls, wide, dd4);
        if (rc) return rc;
        g_last_dd = 1;
        out[0] = dd4[0] + dd4[1];
        out[1] = dd4[2] + dd4[3];
        return PQB200_OK;
    }
    if (rc) return fail(rc, err);
    if (pr.trivial) {
        out[0] = pr.value[0];
        out[1] = pr.value[1];
        return PQB200_OK;
    }
    if (pr.plain_route)  // every digit is binary, no column repeats: a plain permanent
        return plain_multi_device(ls, pr.mat.data(), (size_t)pr.R, (size_t)pr.C, (size_t)pr.C,
                                  (size_t)pr.photons, out);
    const RepPrepared raw = pr;  // the double-double fallback takes the matrix as given
    rep_normalise(pr);
    RepPlan pl;
    int maxr = 1;
    for (int m : pr.r) maxr = std::max(maxr, m);
    const int bstride = maxr + 1;
    rep_plan(pr, ls[0].sm_count * kPlainThreads * pow_min_blocks(pr.C), (unsigned long long)used,
             (unsigned int)pow_len_max(pr.R, pr.C), bstride, pl);
    if ((unsigned long long)used > pl.nchunks) used = (int)pl.nchunks;
    std::vector<double> parts(5 * (size_t)used);
    std::vector<unsigned char> cm(pr.cm.begin(), pr.cm.end());
    static_assert(sizeof(RepStepHost) == 16, "step table layout");
    CUDA_TRY(cudaSetDevice(ls[0].dev));
    CUDA_TRY(cudaEventRecord(ls[0].e0, ls[0].stream));
    for (int d = 0; d < used; d++) {
        DeviceCtx& c = ls[d];
        CUDA_TRY(cudaSetDevice(c.dev));
        if ((rc = grow(&c.rep_base, &c.rep_base_cap, 16 * 2 * (size_t)pr.C))) return rc;
        if ((rc = grow(&c.rep_binom, &c.rep_binom_cap, sizeof(double) * pl.binom.size()))) return rc;
        CUDA_TRY(cudaMemcpyAsync(c.rep_base, pl.base.data(), 16 * 2 * (size_t)pr.C, cudaMemcpyHostToDevice, c.stream));
        CUDA_TRY(cudaMemcpyAsync(c.rep_binom, pl.binom.data(), sizeof(double) * pl.binom.size(), cudaMemcpyHostToDevice, c.stream));
        PowLaunchArgs x;
        x.mat = pr.mat.data();
        x.R = pr.R;
        x.C = pr.C;
        x.cmult = cm.data();
        x.q = pl.q;
        x.len = pl.len;
        x.chunk_begin = pl.nchunks * (unsigned long long)d / (unsigned long long)used;
        x.chunk_end = pl.nchunks * (unsigned long long)(d + 1) / (unsigned long long)used;
        x.sm_count = c.sm_count;
        x.radix = pl.radix;
        x.weight = pl.weight;
        x.table = pl.table.data();
        x.base = c.rep_base;
        x.binom = (const double*)c.rep_binom;
        x.binom_stride = bstride;
        x.low_parity[0] = pl.low_parity[0];
        x.low_parity[1] = pl.low_parity[1];
        for (int k = 0; k < 4; k++) x.phase[k] = pr.phase[k];
        x.partials = c.partials;
        x.counter = c.counter;
        x.out = c.d_out;
        x.stream = c.stream;
        CUDA_TRY(g_pow_launch[pr.C](x));
        CUDA_TRY(cudaMemcpyAsync(c.h_out, c.d_out, 5 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        if (d == 0) CUDA_TRY(cudaEventRecord(c.e1, c.stream));
    }
    double abs_sum = 0.0;
    std::vector<double> p4(4 * (size_t)used);
    for (int d = 0; d < used; d++) {
        DeviceCtx& c = ls[d];
        CUDA_TRY(cudaSetDevice(c.dev));
        CUDA_TRY(cudaStreamSynchronize(c.stream));  // also keeps pl / pr alive until the copies are done
        memcpy(&p4[4 * (size_t)d], c.h_out, 4 * sizeof(double));
        abs_sum += c.h_out[4];
    }
    float ms = 0.f;
    CUDA_TRY(cudaSetDevice(ls[0].dev));
    CUDA_TRY(cudaEventElapsedTime(&ms, ls[0].e0, ls[0].e1));
    g_stats.kernel_ms += ms;
    g_stats.gray_steps += (double)pl.nchunks * (double)pl.len;
    g_stats.launches += (uint64_t)used;
    g_stats.ndev = used;
    // unscaled sum and its conditioning
    cdd tot = cdd_zero();
    for (int d = 0; d < used; d++) {
        cdd x;
        x.re.hi = p4[4 * d + 0];
        x.re.lo = p4[4 * d + 1];
        x.im.hi = p4[4 * d + 2];
        x.im.lo = p4[4 * d + 3];
        cdd_add(tot, x);
    }
    const double mag = hypot(tot.re.hi, tot.im.hi);
    const double bound = mag > 0.0 ? 1.1102230246251565e-16 * (pr.photons + 8) * abs_sum / mag
                                   : (abs_sum > 0.0 ? 1.0 : 0.0);
    g_last_bound = bound;
    static const bool no_dd = getenv("PQB200_NO_DD_FALLBACK") != nullptr;  // measurement aid
    if (bound > kRepeatedDDThreshold && !no_dd) {
        double dd4[4];
        rc = repeated_dd_prepared(ls, raw, dd4);
        if (rc) return rc;
        g_last_dd = 1;
        out[0] = dd4[0] + dd4[1];
        out[1] = dd4[2] + dd4[3];
        return PQB200_OK;
    }
    combine_partials(p4.data(), (size_t)used, (size_t)pr.photons, out);
    return PQB200_OK;
}

int pqb200_last_repeated_bound(double* bound, int* used_dd) {
    if (bound) *bound = g_last_bound;
    if (used_dd) *used_dd = g_last_dd;
    return PQB200_OK;
}

int pqb200_permanent_shard(const double* a, size_t rows, size_t cols, size_t stride, int device,
                           uint64_t shard_index, uint64_t shard_count, double out_dd[4],
                           double* d_out, void* stream) {
    if (shard_count == 0 || shard_index >= shard_count)
        return fail(PQB200_ERR_ARG, "shard_index must be < shard_count");
    if (!out_dd && !d_out) return fail(PQB200_ERR_ARG, "both output pointers are null");
    int rc = check_plain_args(a, rows, cols, stride);
    if (rc) return rc;
    if (rows < 2 || rows >= cols + 2)
        return fail(PQB200_ERR_ARG, "sharded evaluation needs 2 <= rows <= cols+1 (edge cases are not sharded)");
    Lease ls;
    rc = acquire_device(device, ls);
    if (rc) return rc;
    DeviceCtx& c = ls[0];
    PlainPlan pl = plan_plain((int)rows, (int)cols, c.sm_count * kPlainThreads, shard_count);
    const unsigned long long cb = (unsigned long long)((__uint128_t)pl.nchunks * shard_index / shard_count);
    const unsigned long long ce = (unsigned long long)((__uint128_t)pl.nchunks * (shard_index + 1) / shard_count);
    // NULL means the CALLER's legacy default stream when the result stays on the device (d_out): that is
    // the stream a following collective or copy of the caller is ordered against.  With a host result
    // the call synchronises itself and the library's own stream serves.
    cudaStream_t st = stream ? (cudaStream_t)stream : (d_out ? cudaStreamLegacy : c.stream);
    double* dst = d_out ? d_out : c.d_out;
    CUDA_TRY(cudaSetDevice(c.dev));
    CUDA_TRY(cudaEventRecord(c.e0, st));
    rc = launch_plain_range(c, a, rows, cols, stride, pl, cb, ce, dst, st);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(c.e1, st));
    g_stats = {0.0, (double)(ce - cb) * ldexp(1.0, pl.L), 1, 1};
    if (out_dd) {
        CUDA_TRY(cudaMemcpyAsync(c.h_out, dst, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        memcpy(out_dd, c.h_out, 4 * sizeof(double));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, c.e0, c.e1));
        g_stats.kernel_ms = ms;
    } else if (st != c.stream) {
        // work is left in flight on the caller's stream: later users of this slot's reduction
        // workspace are ordered behind it
        if ((rc = mark_async(c, st))) return rc;
    }
    return ls.done(PQB200_OK);
}

int pqb200_permanent_combine(const double* partials_dd, size_t count, size_t rows, double out[2]) {
    if (!partials_dd || !out) return fail(PQB200_ERR_ARG, "null pointer");
    if (rows == 0) return fail(PQB200_ERR_ARG, "rows must be >= 1");
    combine_partials(partials_dd, count, rows, out);
    return PQB200_OK;
}

#include "nccl_capi.inc"
#include "sampler_capi.inc"

struct BatchItem {
    BatchJobHost job;
    BatchJobKind kind;
};

// single-row job: prod_j U[g][p_j] on the host (the reference's rows == 1 branch, hpp:130-140)
static void single_row_value(const double* u, size_t ustride, const BatchJobHost& j, double out[2]) {
    double re = 1.0, im = 0.0;
    for (int c = 0; c < j.N; c++) {
        const size_t g = j.gidx[0], pc = j.pidx[c];
        const size_t off = j.transpose ? (pc * ustride + g) : (g * ustride + pc);
        const double xr = u[2 * off], xi = u[2 * off + 1];
        const double tr = re * xr - im * xi, ti = re * xi + im * xr;
        re = tr;
        im = ti;
    }
    out[0] = re;
    out[1] = im;
}

// Runs the kernel jobs of `items` (all referring to matrix data already resident in c.b_u) on slot
// c and leaves their results in c.b_out[out_slot]; trivial kinds must have been pre-filled by the
// caller in `prefill` (count x complex), which is uploaded first.  Async on c.stream.
static int run_batch_on_device(DeviceCtx& c, const std::vector<BatchItem>& items,
                               const std::vector<double>& prefill, size_t nslots) {
    CUDA_TRY(cudaSetDevice(c.dev));
    if (!c.batch_ready) {
        std::lock_guard<std::mutex> lk(g_sym_mu);  // the binomial table is a per-device __constant__ symbol
        CUDA_TRY(batch_init_device());
        c.batch_ready = true;
    }
    int rc = grow(&c.b_out, &c.b_out_cap, 16 * std::max<size_t>(nslots, 1));
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(c.b_out, prefill.data(), 16 * nslots, cudaMemcpyHostToDevice, c.stream));
    // group the kernel jobs by column bucket (counting sort into one flat array)
    size_t count[64] = {0}, start[64];
    for (const BatchItem& it : items)
        if (it.kind == kJobKernel) count[batch_bucket(it.job.N)]++;
    size_t tot = 0;
    for (int b = 0; b < 64; b++) { start[b] = tot; tot += count[b]; }
    static thread_local std::vector<BatchJobHost> flat;
    flat.resize(tot);
    std::vector<std::pair<int, std::pair<size_t, size_t>>> launches;  // bucket, (offset, count)
    for (int b = 0; b < 64; b++)
        if (count[b]) launches.push_back({b, {start[b], count[b]}});
    for (const BatchItem& it : items)
        if (it.kind == kJobKernel) {
            flat[start[batch_bucket(it.job.N)]++] = it.job;
            g_stats.gray_steps += (double)it.job.len * (double)it.job.nchunks;
        }
    if (flat.empty()) return PQB200_OK;
    rc = grow(&c.b_jobs, &c.b_jobs_cap, sizeof(BatchJobHost) * flat.size());
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(c.b_jobs, flat.data(), sizeof(BatchJobHost) * flat.size(), cudaMemcpyHostToDevice, c.stream));
    CUDA_TRY(cudaStreamSynchronize(c.stream));  // `flat` and `prefill` are host temporaries
    for (auto& l : launches) {
        CUDA_TRY(batch_launch(l.first, c.b_u, (const char*)c.b_jobs + sizeof(BatchJobHost) * l.second.first,
                              (int)l.second.second, c.b_out, c.stream));
        g_stats.launches += 1;
    }
    return PQB200_OK;
}

// FP64 instructions of one repeated permanent with the given occupations (rough: the smaller Gray
// space x 6 x photons); used only to choose the device count of an automatic call
static double repeated_work(size_t rows, size_t cols, const int64_t* row_mult, const int64_t* col_mult) {
    double sr = 1.0, sc = 1.0, n = 0.0;
    for (size_t i = 0; i < rows; i++) if (row_mult[i] > 0) { sr *= (double)(row_mult[i] + 1); n += (double)row_mult[i]; }
    for (size_t j = 0; j < cols; j++) if (col_mult[j] > 0) sc *= (double)(col_mult[j] + 1);
    return std::min(sr, sc) * 6.0 * n;
}

int pqb200_permanent_repeated(const double* a, size_t rows, size_t cols, size_t stride,
                              const int64_t* row_mult, const int64_t* col_mult, int ndev,
                              double out[2]) {
    if (!out) return fail(PQB200_ERR_ARG, "output pointer is null");
    int rc = check_repeated_args(a, rows, cols, stride, row_mult, col_mult);
    if (rc) return rc;
    Lease ls;
    rc = acquire(ndev, repeated_work(rows, cols, row_mult, col_mult), ls);
    if (rc) return rc;
    g_stats = {0.0, 0.0, 0, ls.size()};
    return ls.done(repeated_locked(ls, a, rows, cols, stride, row_mult, col_mult, out));
}

int pqb200_permanent_repeated_batch(const double* a, size_t rows, size_t cols, size_t stride,
                                    const int64_t* row_mults, const int64_t* col_mults,
                                    size_t count, int ndev, double* out) {
    if (count && !out) return fail(PQB200_ERR_ARG, "output pointer is null");
    if (count == 0) return PQB200_OK;
    std::vector<BatchItem> items(count);
    std::vector<double> prefill(2 * count, 0.0);
    std::vector<size_t> too_big;
    double work = 0.0;
    for (size_t k = 0; k < count; k++) {
        int rc = check_repeated_args(a, rows, cols, stride, row_mults + k * rows, col_mults + k * cols);
        if (rc) return rc;
        std::string err;
        BatchItem& it = items[k];
        memset(&it.job, 0, sizeof(it.job));
        rc = make_batch_job(rows, cols, row_mults + k * rows, col_mults + k * cols, it.job, it.kind, err);
        if (rc) return fail(rc, err);
        it.job.out_slot = (int)k;
        it.job.ustride = (int)stride;
        it.job.uoff = 0;
        if (it.kind == kJobTrivialOne) prefill[2 * k] = 1.0;
        else if (it.kind == kJobSingleRow) single_row_value(a, stride, it.job, &prefill[2 * k]);
        else if (it.kind == kJobTooBig) {
            too_big.push_back(k);
            work += repeated_work(rows, cols, row_mults + k * rows, col_mults + k * cols);
        } else work += (double)it.job.len * (double)it.job.nchunks * 6.0 * it.job.N;
    }
    Lease ls;
    int rc = acquire(ndev, work, ls);
    if (rc) return rc;
    int used = ls.size();
    g_stats = {0.0, 0.0, 0, used};
    // shard the items over the devices by contiguous ranges
    used = (int)std::min<size_t>((size_t)used, count);
    std::vector<std::vector<BatchItem>> per_dev((size_t)used);
    for (size_t k = 0; k < count; k++) per_dev[k * (size_t)used / count].push_back(items[k]);
    CUDA_TRY(cudaSetDevice(ls[0].dev));
    CUDA_TRY(cudaEventRecord(ls[0].e0, ls[0].stream));
    for (int d = 0; d < used; d++) {
        DeviceCtx& c = ls[d];
        CUDA_TRY(cudaSetDevice(c.dev));
        if (rows && cols) {
            rc = grow(&c.b_u, &c.b_u_cap, 16 * rows * stride);
            if (rc) return rc;
            CUDA_TRY(cudaMemcpyAsync(c.b_u, a, 16 * ((rows - 1) * stride + cols), cudaMemcpyHostToDevice, c.stream));
        }
        rc = run_batch_on_device(c, per_dev[(size_t)d], prefill, count);
        if (rc) return rc;
        if (d == 0) CUDA_TRY(cudaEventRecord(c.e1, c.stream));
    }
    std::vector<double> res(2 * count);
    for (int d = 0; d < used; d++) {
        DeviceCtx& c = ls[d];
        CUDA_TRY(cudaSetDevice(c.dev));
        CUDA_TRY(cudaMemcpyAsync(res.data(), c.b_out, 16 * count, cudaMemcpyDeviceToHost, c.stream));
        CUDA_TRY(cudaStreamSynchronize(c.stream));
        for (const BatchItem& it : per_dev[(size_t)d]) {
            out[2 * it.job.out_slot] = res[2 * it.job.out_slot];
            out[2 * it.job.out_slot + 1] = res[2 * it.job.out_slot + 1];
        }
    }
    float ms = 0.f;
    CUDA_TRY(cudaSetDevice(ls[0].dev));
    CUDA_TRY(cudaEventElapsedTime(&ms, ls[0].e0, ls[0].e1));
    g_stats.kernel_ms = ms;
    // jobs whose Gray space is too large for one CTA use the grid-wide kernels, one at a time
    for (size_t k : too_big) {
        rc = repeated_locked(ls, a, rows, cols, stride, row_mults + k * rows, col_mults + k * cols, out + 2 * k);
        if (rc) return rc;
    }
    return ls.done(PQB200_OK);
}

int pqb200_permanent_batch(const double* const* a, const size_t* rows, const size_t* cols,
                           const size_t* strides, size_t count, int ndev, double* out) {
    if (!out || (count && (!a || !rows || !cols || !strides))) return fail(PQB200_ERR_ARG, "null pointer");
    if (count == 0) return PQB200_OK;
    std::vector<size_t> single;  // matrices that do not fit one CTA (or are not square): grid-wide path
    {
        std::vector<BatchItem> items;
        std::vector<double> prefill(2 * count, 0.0), packed;
        const int64_t ones[kBatchColsMax] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                             1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
        for (size_t k = 0; k < count; k++) {
            const size_t n = rows[k];
            if (n != cols[k] || n < 2 || n > 22) {
                single.push_back(k);
                continue;
            }
            int rc = check_plain_args(a[k], n, n, strides[k]);
            if (rc) return rc;
            BatchItem it;
            memset(&it.job, 0, sizeof(it.job));
            std::string err;
            rc = make_batch_job(n, n, ones, ones, it.job, it.kind, err);
            if (rc) return fail(rc, err);
            if (it.kind != kJobKernel) {
                single.push_back(k);
                continue;
            }
            it.job.out_slot = (int)k;
            it.job.ustride = (int)n;
            it.job.uoff = packed.size() / 2;
            for (size_t i = 0; i < n; i++)
                packed.insert(packed.end(), a[k] + 2 * i * strides[k], a[k] + 2 * (i * strides[k] + n));
            items.push_back(it);
        }
        g_stats = {0.0, 0.0, 0, 1};
        if (!items.empty()) {
            Lease ls;
            int rc = acquire(ndev <= 0 ? 0 : 1, 0.0, ls);  // small matrices: one device is plenty (launch-latency bound)
            if (rc) return rc;
            DeviceCtx& c = ls[0];
            CUDA_TRY(cudaSetDevice(c.dev));
            rc = grow(&c.b_u, &c.b_u_cap, sizeof(double) * packed.size());
            if (rc) return rc;
            CUDA_TRY(cudaEventRecord(c.e0, c.stream));
            CUDA_TRY(cudaMemcpyAsync(c.b_u, packed.data(), sizeof(double) * packed.size(), cudaMemcpyHostToDevice, c.stream));
            rc = run_batch_on_device(c, items, prefill, count);
            if (rc) return rc;
            CUDA_TRY(cudaEventRecord(c.e1, c.stream));
            std::vector<double> res(2 * count);
            CUDA_TRY(cudaMemcpyAsync(res.data(), c.b_out, 16 * count, cudaMemcpyDeviceToHost, c.stream));
            CUDA_TRY(cudaStreamSynchronize(c.stream));
            for (const BatchItem& it : items) {
                out[2 * it.job.out_slot] = res[2 * it.job.out_slot];
                out[2 * it.job.out_slot + 1] = res[2 * it.job.out_slot + 1];
            }
            float ms = 0.f;
            CUDA_TRY(cudaEventElapsedTime(&ms, c.e0, c.e1));
            g_stats.kernel_ms = ms;
            ls.done(PQB200_OK);
        }
    }
    pqb200_stats acc = g_stats;
    for (size_t k : single) {
        int rc = pqb200_permanent(a[k], rows[k], cols[k], strides[k], ndev, out + 2 * k);
        if (rc) return rc;
        acc.kernel_ms += g_stats.kernel_ms;
        acc.gray_steps += g_stats.gray_steps;
        acc.launches += g_stats.launches;
        acc.ndev = std::max(acc.ndev, g_stats.ndev);
    }
    g_stats = acc;
    return PQB200_OK;
}

// pmf rows for shots [s0, s1) on slot c; async on c.stream, results left in c.b_pmf
static int pmf_layer_on_device(Lease& ls, DeviceCtx& c, const double* u, size_t m, size_t stride,
                               const int64_t* samples, const int64_t* inputs, size_t s0, size_t s1) {
    const size_t S = s1 - s0;
    const double t_begin = now_ms();
    static thread_local std::vector<BatchItem> items;
    static thread_local std::vector<double> prefill, weight;
    items.clear();
    prefill.assign(2 * S * m, 0.0);
    weight.assign(S * m, 0.0);
    // job construction is pure index work on disjoint shots: spread it over a few host threads
    struct TooBig { size_t shot, mode, slot; };
    const int nthreads = (int)std::max<size_t>(1, std::min<size_t>({(size_t)8, (size_t)std::thread::hardware_concurrency(), S / 64}));
    std::vector<std::vector<BatchItem>> t_items((size_t)nthreads);
    std::vector<std::vector<TooBig>> t_big((size_t)nthreads);
    std::vector<int> t_rc((size_t)nthreads, 0);
    std::vector<const char*> t_err((size_t)nthreads, nullptr);
    // NB: `prefill` / `weight` are thread_local; worker threads must go through these pointers
    double* const prefill_p = prefill.data();
    double* const weight_p = weight.data();
    auto build = [&, prefill_p, weight_p](int t) {
        const size_t b0 = s0 + S * (size_t)t / (size_t)nthreads, b1 = s0 + S * (size_t)(t + 1) / (size_t)nthreads;
        std::vector<int64_t> cols(m);
        BatchSide rs, cs;
        std::vector<BatchItem>& out = t_items[(size_t)t];
        out.reserve((b1 - b0) * 24);
        for (size_t s = b0; s < b1; s++) {
            const int64_t* sample = samples + s * m;
            const int64_t* input = inputs + s * m;
            analyze_side(m, sample, rs);  // the row side is shared by all sub-permanents of the shot
            memcpy(cols.data(), input, m * sizeof(int64_t));
            for (size_t i = 0; i < m; i++) {
                if (input[i] <= 0) continue;
                const size_t slot = (s - s0) * m + i;
                weight_p[slot] = (double)input[i];
                cols[i]--;  // input_loc = current_input - e_i (samplingHelperFunctions.cpp:139-143)
                analyze_side(m, cols.data(), cs);
                cols[i]++;
                BatchItem it;
                memset(&it.job, 0, sizeof(it.job));
                const char* err = nullptr;
                const int rc = make_batch_job_sides(rs, cs, it.job, it.kind, &err);
                if (rc) {
                    t_rc[(size_t)t] = rc;
                    t_err[(size_t)t] = err;
                    return;
                }
                it.job.out_slot = (int)slot;
                it.job.ustride = (int)stride;
                it.job.uoff = 0;
                if (it.kind == kJobKernel) out.push_back(it);
                else if (it.kind == kJobTrivialOne) prefill_p[2 * slot] = 1.0;
                else if (it.kind == kJobSingleRow) single_row_value(u, stride, it.job, &prefill_p[2 * slot]);
                else t_big[(size_t)t].push_back({s, i, slot});
            }
        }
    };
    if (nthreads == 1) {
        build(0);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; t++) pool.emplace_back(build, t);
        for (auto& th : pool) th.join();
    }
    for (int t = 0; t < nthreads; t++) {
        if (t_rc[(size_t)t]) return fail(t_rc[(size_t)t], t_err[(size_t)t]);
        items.insert(items.end(), t_items[(size_t)t].begin(), t_items[(size_t)t].end());
        for (const TooBig& tb : t_big[(size_t)t]) {
            // Gray space too large for one CTA: grid-wide kernels on this slot, synchronously
            std::vector<int64_t> cols(inputs + tb.shot * m, inputs + (tb.shot + 1) * m);
            cols[tb.mode]--;
            double v[2];
            Lease one;  // a view of this slot only (already locked by `ls`)
            one.slots.push_back(&c);
            const int rc = repeated_locked(one, u, m, m, stride, samples + tb.shot * m, cols.data(), v);
            one.slots.clear();
            if (rc) return rc;
            prefill[2 * tb.slot] = v[0];
            prefill[2 * tb.slot + 1] = v[1];
        }
    }
    (void)ls;
    const double t_jobs = now_ms();
    CUDA_TRY(cudaSetDevice(c.dev));
    int rc = grow(&c.b_u, &c.b_u_cap, 16 * m * stride);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(c.b_u, u, 16 * ((m - 1) * stride + m), cudaMemcpyHostToDevice, c.stream));
    rc = grow(&c.b_weight, &c.b_weight_cap, sizeof(double) * S * m);
    if (rc) return rc;
    rc = grow(&c.b_pmf, &c.b_pmf_cap, sizeof(double) * S * m);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(c.b_weight, weight.data(), sizeof(double) * S * m, cudaMemcpyHostToDevice, c.stream));
    rc = run_batch_on_device(c, items, prefill, S * m);  // synchronises the uploads of the host temporaries
    if (rc) return rc;
    CUDA_TRY(pmf_combine_launch(c.b_u, (int)stride, (int)m, c.b_out, (const double*)c.b_weight,
                                (double*)c.b_pmf, (int)S, c.stream));
    g_stats.launches += 1;
    if (trace_on())
        fprintf(stderr, "[pqb200] pmf layer dev %d: %zu shots, %zu kernel jobs; build %.2f ms, upload+launch %.2f ms\n",
                c.dev, S, items.size(), t_jobs - t_begin, now_ms() - t_jobs);
    return PQB200_OK;
}

static int pmf_layer_locked(Lease& ls, const double* u, size_t m, size_t stride, const int64_t* samples,
                            const int64_t* inputs, size_t S, double* pmf) {
    const int used = (int)std::min<size_t>((size_t)ls.size(), S);
    const double t_layer = now_ms();
    CUDA_TRY(cudaSetDevice(ls[0].dev));
    CUDA_TRY(cudaEventRecord(ls[0].e0, ls[0].stream));
    for (int d = 0; d < used; d++) {
        const size_t s0 = S * (size_t)d / (size_t)used, s1 = S * (size_t)(d + 1) / (size_t)used;
        int rc = pmf_layer_on_device(ls, ls[d], u, m, stride, samples, inputs, s0, s1);
        if (rc) return rc;
        if (d == 0) CUDA_TRY(cudaEventRecord(ls[0].e1, ls[0].stream));
    }
    for (int d = 0; d < used; d++) {
        DeviceCtx& c = ls[d];
        const size_t s0 = S * (size_t)d / (size_t)used, s1 = S * (size_t)(d + 1) / (size_t)used;
        CUDA_TRY(cudaSetDevice(c.dev));
        CUDA_TRY(cudaMemcpyAsync(pmf + s0 * m, c.b_pmf, sizeof(double) * (s1 - s0) * m, cudaMemcpyDeviceToHost, c.stream));
        CUDA_TRY(cudaStreamSynchronize(c.stream));
    }
    float ms = 0.f;
    CUDA_TRY(cudaSetDevice(ls[0].dev));
    CUDA_TRY(cudaEventElapsedTime(&ms, ls[0].e0, ls[0].e1));
    g_stats.kernel_ms += ms;
    g_stats.ndev = used;
    if (trace_on()) fprintf(stderr, "[pqb200] pmf layer total %.2f ms (device-0 stream %.2f ms)\n", now_ms() - t_layer, ms);
    return PQB200_OK;
}

static int check_pmf_args(const double* u, size_t m, size_t stride, const int64_t* samples,
                          const int64_t* inputs, double* pmf) {
    if (!u || !samples || !inputs || !pmf) return fail(PQB200_ERR_ARG, "null pointer");
    if (m == 0 || m > 255) return fail(PQB200_ERR_ARG, "number of modes must be in 1..255");
    if (stride < m) return fail(PQB200_ERR_ARG, "stride is smaller than the number of modes");
    return PQB200_OK;
}

static double max_photons(const int64_t* inputs, size_t S, size_t m) {
    int64_t mx = 0;
    for (size_t s = 0; s < S; s++) {
        int64_t n = 0;
        for (size_t i = 0; i < m; i++) n += inputs[s * m + i];
        mx = std::max(mx, n);
    }
    return (double)mx;
}

int pqb200_pmf_layer(const double* u, size_t m, size_t stride, const int64_t* samples,
                     const int64_t* inputs, size_t S, int ndev, double* pmf) {
    int rc = check_pmf_args(u, m, stride, samples, inputs, pmf);
    if (rc) return rc;
    if (S == 0) return PQB200_OK;
    Lease ls;
    rc = acquire(ndev, laplace_work((double)S, max_photons(inputs, S, m)), ls);
    if (rc) return rc;
    g_stats = {0.0, 0.0, 0, ls.size()};
    // all sub-permanents of a shot in one Gray sweep (laplace_sampler.cuh); shots beyond that
    // kernel's photon limit fall through to one batched job per sub-permanent
    rc = pmf_layer_laplace_locked(ls, u, m, stride, samples, inputs, S, pmf);
    if (rc != -1) return ls.done(rc);
    return ls.done(pmf_layer_locked(ls, u, m, stride, samples, inputs, S, pmf));
}

// The same layer with one batched kernel job per sub-permanent (k repeated permanents per shot, as
// compute_pmf itself does): kept as the cross-check of the single-sweep kernel and for > 40 photons.
int pqb200_pmf_layer_batched(const double* u, size_t m, size_t stride, const int64_t* samples,
                             const int64_t* inputs, size_t S, int ndev, double* pmf) {
    int rc = check_pmf_args(u, m, stride, samples, inputs, pmf);
    if (rc) return rc;
    if (S == 0) return PQB200_OK;
    Lease ls;
    const double ph = max_photons(inputs, S, m);
    rc = acquire(ndev, ph * laplace_work((double)S, ph), ls);
    if (rc) return rc;
    g_stats = {0.0, 0.0, 0, ls.size()};
    return ls.done(pmf_layer_locked(ls, u, m, stride, samples, inputs, S, pmf));
}

}  // extern "C"
