// rng_replay.cpp -- CPU check of csrc/rng_replay.h (built and run by tests/test_host_logic_cpu.py):
// the block generator against std::mt19937 (the reference's generator, samplingHelperFunctions.cpp:58-74),
// the multiplicative remainder against %, and draw_photon_loop against a plain restatement of the
// reference's per-step draws.  Prints "ok" and the replay speed.
#include <chrono>
#include <cstdio>
#include <random>
#include <vector>

#include "../../piquassoboost_b200/csrc/rng_replay.h"

using pqb200::ReplayRng;

static int fail(const char* what, long long at) {
    printf("FAIL %s at %lld\n", what, at);
    return 1;
}

// the reference's helpers, verbatim in meaning, on a std::mt19937
static double ref_double(std::mt19937& g) { return static_cast<double>(g()) / static_cast<double>(g.max()); }
static size_t ref_index(std::mt19937& g, size_t upper) { return upper == 0 ? 0 : g() % upper; }

static void ref_photon_loop(std::mt19937& g, std::vector<uint8_t> particles, int64_t steps, std::vector<uint8_t>& ins,
                            std::vector<double>& rnd) {
    ins.clear();
    rnd.clear();
    while (!particles.empty() && (steps < 0 || (int64_t)ins.size() < steps)) {
        const size_t pick = ref_index(g, particles.size());
        ins.push_back(particles[pick]);
        particles.erase(particles.begin() + (long)pick);
        rnd.push_back(ref_double(g));
    }
}

int main() {
    // 1. the raw stream, default seed and explicit seeds (64-bit seeds are taken mod 2^32), across many refills
    {
        ReplayRng r;
        std::mt19937 g;
        for (int i = 0; i < 2000; i++)
            if (r.next() != (uint32_t)g()) return fail("default-seeded stream", i);
    }
    const uint64_t seeds[] = {0ull, 1ull, 5ull, 7ull, 1234567ull, 0xffffffffull, 0x100000005ull, 0xdeadbeefcafef00dull};
    for (uint64_t sd : seeds) {
        ReplayRng r;
        std::mt19937 g;
        r.seed(sd);
        g.seed(static_cast<std::mt19937::result_type>(sd));
        for (int i = 0; i < 100000; i++)
            if (r.next() != (uint32_t)g()) return fail("seeded stream", i);
        // mixed consumption: runs of every length (block boundaries included), doubles and indices
        for (int rep = 0; rep < 3000; rep++) {
            const size_t cnt = 1 + (size_t)rep % ReplayRng::kRun;
            const uint32_t* p = r.take(cnt);
            for (size_t i = 0; i < cnt; i++)
                if (p[i] != (uint32_t)g()) return fail("take()", rep);
            if (r.double_0_1() != ref_double(g)) return fail("double_0_1", rep);
            const size_t upper = rep % 7 == 0 ? 0 : (size_t)(1 + (rep * 37) % 300);
            if (r.index(upper) != ref_index(g, upper)) return fail("index", rep);
        }
        if (r.index(0x100000000ull) != ref_index(g, 0x100000000ull)) return fail("index, 64-bit upper", 0);
        if (r.index(0xffffffffull) != ref_index(g, 0xffffffffull)) return fail("index, upper 2^32-1", 0);
    }
    // 2. the remainder by multiplication, every table divisor against a spread of values and the extremes
    for (uint32_t d = 1; d <= 300; d++) {
        uint32_t x = 0x9e3779b9u * d;
        for (int i = 0; i < 20000; i++) {
            if (ReplayRng::mod(x, d) != x % d) return fail("mod", d);
            x = x * 1664525u + 1013904223u;
        }
        const uint32_t edge[] = {0u, 1u, d - 1, d, d + 1, 0x7fffffffu, 0x80000000u, 0xfffffffeu, 0xffffffffu};
        for (uint32_t e : edge)
            if (ReplayRng::mod(e, d) != e % d) return fail("mod edge", d);
    }
    // 3. photon loops: full and truncated, repeated modes, 0..40 photons, stream position afterwards
    {
        ReplayRng r;
        std::mt19937 g;
        r.seed(99);
        g.seed(99);
        std::vector<uint8_t> ins_ref;
        std::vector<double> rnd_ref;
        for (int shot = 0; shot < 20000; shot++) {
            const size_t count = (size_t)shot % 41;
            std::vector<uint8_t> particles(count);
            for (size_t i = 0; i < count; i++) particles[i] = (uint8_t)((i * 3 + (size_t)shot) % 7 + i / 2);
            const int64_t steps = shot % 3 == 0 ? -1 : (int64_t)((size_t)shot % (count + 2));
            uint8_t ins[40], ins_s[40];
            double rnd[40], rnd_s[40];
            ReplayRng copy = r;  // the portable form on the same outputs (the dispatch may have taken the BMI2 form)
            pqb200::draw_photon_loop(r, particles.data(), count, steps, ins, rnd);
            const size_t todo = steps < 0 || (size_t)steps > count ? count : (size_t)steps;
            if (todo) pqb200::draw_photon_loop_shift(copy.take(2 * todo), particles.data(), count, todo, ins_s, rnd_s);
            for (size_t k = 0; k < todo; k++)
                if (ins[k] != ins_s[k] || rnd[k] != rnd_s[k]) return fail("portable photon loop", shot);
            ref_photon_loop(g, particles, steps, ins_ref, rnd_ref);
            for (size_t k = 0; k < ins_ref.size(); k++)
                if (ins[k] != ins_ref[k] || rnd[k] != rnd_ref[k]) return fail("photon loop", shot);
            if (r.next() != (uint32_t)g()) return fail("stream position after a photon loop", shot);
        }
    }
#if PQB200_REPLAY_BMI2
    const int has_bmi2 = pqb200::replay_has_bmi2();
#else
    const int has_bmi2 = 0;
#endif
    // 4. speed: C5-shaped shots (20 photons), this header against the plain std::mt19937 loop
    {
        const size_t shots = 100000, n = 20;
        std::vector<uint8_t> particles(n), ins(shots * n);
        for (size_t i = 0; i < n; i++) particles[i] = (uint8_t)i;
        std::vector<double> rnd(shots * n);
        ReplayRng r;
        r.seed(5);
        auto t0 = std::chrono::steady_clock::now();
        for (size_t s = 0; s < shots; s++) pqb200::draw_photon_loop(r, particles.data(), n, -1, ins.data() + s * n, rnd.data() + s * n);
        auto t1 = std::chrono::steady_clock::now();
        std::mt19937 g;
        g.seed(5);
        std::vector<uint8_t> a;
        std::vector<double> b;
        double chk = 0.0;
        for (size_t s = 0; s < shots; s++) {
            ref_photon_loop(g, particles, -1, a, b);
            chk += b[7] + a[3];
        }
        auto t2 = std::chrono::steady_clock::now();
        printf("ok bmi2=%d replay %.0f ns/shot, plain std::mt19937 loop %.0f ns/shot (check %g %g)\n", has_bmi2,
               std::chrono::duration<double, std::nano>(t1 - t0).count() / shots,
               std::chrono::duration<double, std::nano>(t2 - t1).count() / shots, rnd[12345], chk);
    }
    return 0;
}
