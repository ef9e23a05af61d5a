// rng_replay.h -- host-only replay of the reference's random stream for the B samplers (no CUDA in here:
// tests/cpp/rng_replay.cpp compiles it with g++ and checks it against std::mt19937 on the CPU).
//
// The reference draws everything from ONE process-global std::mt19937 (samplingHelperFunctions.cpp:58-74):
//     random_double_0_1() = double(rng()) / double(rng.max())                (:64-66)
//     random_index(upper) = rng() % upper          (0 for upper == 0)        (:68-74)
// and a shot's photon loop consumes, per step, one random_index (which photon enters,
// update_input_by_single_photon :377-406) and one random_double_0_1 (sample_from_pmf :355).  None of it
// depends on the pmf values, so the host replays the stream shot after shot and the GPU does the rest.  The
// replay is serial by definition (one stream), which makes its speed the floor of a many-GPU sampler call:
// hence a block generator (624 outputs per refill, loops the compiler vectorises), remainders by
// multiplication (Lemire's exact 32-bit fastmod, table of magic numbers), outputs taken in runs, and the
// pick of the k-th photon still outside by PDEP/TZCNT where the CPU has BMI2.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

namespace pqb200 {

/// The MT19937 generator of std::mt19937, output for output (seeding as mersenne_twister_engine::seed(value)).
class ReplayRng {
public:
    static constexpr int kN = 624, kM = 397;
    ReplayRng() { seed(5489u); }  // std::mt19937's default seed

    void seed(uint64_t value) {
        mt_[0] = (uint32_t)value;  // value mod 2^32
        for (int i = 1; i < kN; i++) mt_[i] = 1812433253u * (mt_[i - 1] ^ (mt_[i - 1] >> 30)) + (uint32_t)i;
        pos_ = kN;
    }

    /// next 32-bit output
    inline uint32_t next() {
        if (pos_ >= kN) refill();
        return out_[pos_++];
    }

    /// `cnt` (<= kRun) consecutive outputs: a pointer into the block when they are all there, else a copy
    static constexpr int kRun = 96;
    inline const uint32_t* take(size_t cnt) {
        if (pos_ + (int)cnt <= kN) {
            const uint32_t* p = out_ + pos_;
            pos_ += (int)cnt;
            return p;
        }
        for (size_t i = 0; i < cnt; i++) run_[i] = next();
        return run_;
    }

    static inline double to_double_0_1(uint32_t raw) { return (double)raw / 4294967295.0; }  // rng.max() = 2^32 - 1
    inline double double_0_1() { return to_double_0_1(next()); }

    /// raw % upper for 32-bit values, exactly (D. Lemire, "Faster remainder by direct computation", 2019)
    static inline uint32_t mod(uint32_t raw, uint32_t upper) {
        if (upper < kMagic) {
            const uint64_t low = magic()[upper] * raw;
            return (uint32_t)(((unsigned __int128)low * upper) >> 64);
        }
        return raw % upper;
    }
    inline size_t index(size_t upper) {
        if (upper == 0) return 0;  // the reference returns 0 WITHOUT consuming an output
        // the outputs are 32-bit values: for upper >= 2^32 the 64-bit remainder is the value itself
        const uint32_t raw = next();
        return upper > 0xffffffffull ? (size_t)raw : (size_t)mod(raw, (uint32_t)upper);
    }

private:
    static constexpr uint32_t kMagic = 257;
    static const uint64_t* magic() {
        static const struct Table {
            uint64_t m[kMagic];
            Table() {
                m[0] = 0;
                for (uint32_t d = 1; d < kMagic; d++) m[d] = 0xffffffffffffffffull / d + 1;  // d = 1: wraps to 0
            }
        } t;
        return t.m;
    }

    void refill() {
        constexpr uint32_t UP = 0x80000000u, LO = 0x7fffffffu, A = 0x9908b0dfu;
        uint32_t* mt = mt_;
        for (int k = 0; k < kN - kM; k++) {
            const uint32_t y = (mt[k] & UP) | (mt[k + 1] & LO);
            mt[k] = mt[k + kM] ^ (y >> 1) ^ ((0u - (y & 1u)) & A);
        }
        for (int k = kN - kM; k < kN - 1; k++) {
            const uint32_t y = (mt[k] & UP) | (mt[k + 1] & LO);
            mt[k] = mt[k + kM - kN] ^ (y >> 1) ^ ((0u - (y & 1u)) & A);
        }
        {
            const uint32_t y = (mt[kN - 1] & UP) | (mt[0] & LO);
            mt[kN - 1] = mt[kM - 1] ^ (y >> 1) ^ ((0u - (y & 1u)) & A);
        }
        for (int k = 0; k < kN; k++) {  // tempering
            uint32_t z = mt[k];
            z ^= z >> 11;
            z ^= (z << 7) & 0x9d2c5680u;
            z ^= (z << 15) & 0xefc60000u;
            z ^= z >> 18;
            out_[k] = z;
        }
        pos_ = 0;
    }

    uint32_t mt_[kN], out_[kN], run_[kRun];
    int pos_;
};

constexpr int kReplayPhotonsMax = 40;  // photons per shot (= kLapPhotonsMax, laplace_host.h)

/// One shot's photon loop as the reference's generator sees it: `particles` (count <= 40) is the working
/// input state (mode of every photon still outside, in mode order); per step random_index picks one
/// (update_input_by_single_photon, samplingHelperFunctions.cpp:377-406) and random_double_0_1 is drawn for
/// sample_from_pmf (:355).  `steps` photons are inserted (all of them when negative).  ins[k] = input mode
/// of the k-th inserted photon, rnd[k] = the uniform draw of step k.
///
/// Portable form: the working list is a byte array, removing an entry is a short shift.
inline void draw_photon_loop_shift(const uint32_t* raw, const uint8_t* particles, size_t left, size_t todo,
                                   uint8_t* ins, double* rnd) {
    uint8_t work[kReplayPhotonsMax + 16] = {0};
    memcpy(work, particles, left);
    for (size_t k = 0; k < todo; k++) {
        const size_t pick = ReplayRng::mod(raw[2 * k], (uint32_t)left);
        ins[k] = work[pick];
        for (size_t w = pick; w + 1 < left; w += 8) {  // work[w .. w+8) <- work[w+1 .. w+9): ascending, so no overlap
            uint64_t v;
            memcpy(&v, work + w + 1, 8);
            memcpy(work + w, &v, 8);
        }
        left--;
        rnd[k] = ReplayRng::to_double_0_1(raw[2 * k + 1]);
    }
}

#if defined(__x86_64__) && defined(__GNUC__)
#define PQB200_REPLAY_BMI2 1
/// BMI2 form: the photons still outside are the set bits of a mask over the ORIGINAL list, and the pick-th of
/// them is found with one PDEP + TZCNT: no memory traffic on the step-to-step dependency chain (3x faster).
__attribute__((target("bmi2"))) inline void draw_photon_loop_bmi2(const uint32_t* raw, const uint8_t* particles,
                                                                  size_t left, size_t todo, uint8_t* ins,
                                                                  double* rnd) {
    unsigned long long mask = (1ull << left) - 1ull;  // left <= 40
    for (size_t k = 0; k < todo; k++) {
        const uint32_t pick = ReplayRng::mod(raw[2 * k], (uint32_t)left);
        const unsigned long long bit = __builtin_ia32_pdep_di(1ull << pick, mask);
        ins[k] = particles[__builtin_ctzll(bit)];
        mask ^= bit;
        left--;
        rnd[k] = ReplayRng::to_double_0_1(raw[2 * k + 1]);
    }
}
inline bool replay_has_bmi2() {
    static const bool has = __builtin_cpu_supports("bmi2") != 0;
    return has;
}
#else
#define PQB200_REPLAY_BMI2 0
#endif

inline void draw_photon_loop(ReplayRng& rng, const uint8_t* particles, size_t count, int64_t steps, uint8_t* ins,
                             double* rnd) {
    const size_t left = count < (size_t)kReplayPhotonsMax ? count : (size_t)kReplayPhotonsMax;
    size_t todo = left;
    if (steps >= 0 && (size_t)steps < todo) todo = (size_t)steps;
    if (todo == 0) return;
    const uint32_t* raw = rng.take(2 * todo);  // two outputs per step, in the reference's order
#if PQB200_REPLAY_BMI2
    if (replay_has_bmi2()) return draw_photon_loop_bmi2(raw, particles, left, todo, ins, rnd);
#endif
    draw_photon_loop_shift(raw, particles, left, todo, ins, rnd);
}

}  // namespace pqb200
