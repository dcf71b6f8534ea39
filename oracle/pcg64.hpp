// ORACLE — test infrastructure only.  Nothing under oracle/ is part of the product path.
//
// Restatement of the random-number arithmetic the reference takes from crates that are NOT
// vendored under /root/reference (Cargo.lock:325-366):
//   rand_core 0.6.2  SeedableRng::seed_from_u64      (called at src/rngator.rs:30)
//   rand_pcg  0.3.0  Pcg64 = Lcg128Xsl64             (src/rngator.rs:28)
//   rand      0.8.3  Rng::gen_range for Range<f64> and Range<usize>
//                    (src/vec.rs:16,47  src/bhv.rs:127  src/textures.rs:142 ...)
// The published algorithms are restated from the crates' documentation/known behaviour and pinned by
// upstream known-answer vectors (tests/test_oracle_rng.py) and by an independent
// pure-Python big-integer restatement (tests/golden/make_golden.py).
#pragma once
#include <cstdint>
#include <cstring>

namespace orc {

typedef unsigned __int128 u128;

struct Pcg64 {
    u128 state;
    u128 increment;
    uint64_t draws = 0;  // number of next_u64 calls (instrumentation; not part of the algorithm)

    static constexpr u128 multiplier() {
        return ((u128)0x2360ED051FC65DA4ULL << 64) | (u128)0x4385DF649FCCF645ULL;
    }

    void step() { state = state * multiplier() + increment; }

    // Lcg128Xsl64::from_state_incr
    static Pcg64 from_state_incr(u128 st, u128 incr) {
        Pcg64 p;
        p.state = st + incr;
        p.increment = incr;
        p.step();
        return p;
    }

    // Lcg128Xsl64::new(state, stream)
    static Pcg64 from_state_stream(u128 st, u128 stream) { return from_state_incr(st, (stream << 1) | 1); }

    // Lcg128Xsl64::from_seed: four little-endian u64
    static Pcg64 from_seed(const uint8_t seed[32]) {
        uint64_t s[4];
        for (int i = 0; i < 4; i++) {
            uint64_t v = 0;
            for (int b = 7; b >= 0; b--) v = (v << 8) | seed[8 * i + b];
            s[i] = v;
        }
        u128 st = (u128)s[0] | ((u128)s[1] << 64);
        u128 inc = (u128)s[2] | ((u128)s[3] << 64);
        return from_state_incr(st, inc | 1);
    }

    // rand_core::SeedableRng::seed_from_u64: PCG32 expands the u64 into the 32-byte seed
    static Pcg64 seed_from_u64(uint64_t st) {
        uint8_t seed[32];
        for (int i = 0; i < 8; i++) {
            st = st * 6364136223846793005ULL + 11634580027462260723ULL;
            uint32_t xorshifted = (uint32_t)(((st >> 18) ^ st) >> 27);
            uint32_t rot = (uint32_t)(st >> 59);
            uint32_t x = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
            seed[4 * i + 0] = (uint8_t)(x);
            seed[4 * i + 1] = (uint8_t)(x >> 8);
            seed[4 * i + 2] = (uint8_t)(x >> 16);
            seed[4 * i + 3] = (uint8_t)(x >> 24);
        }
        return from_seed(seed);
    }

    uint64_t next_u64() {
        draws++;
        step();
        uint32_t rot = (uint32_t)(state >> 122);
        uint64_t xsl = (uint64_t)(state >> 64) ^ (uint64_t)state;
        return (xsl >> rot) | (xsl << ((64 - rot) & 63));
    }

    // rand 0.8 UniformFloat<f64>::sample_single: 52 mantissa bits into [1,2), minus 1, scale, add.
    // Multiply and add are separate roundings (no FMA): compile with -ffp-contract=off.
    double gen_range_f64(double low, double high) {
        double scale = high - low;
        for (;;) {
            uint64_t bits = (next_u64() >> 12) | 0x3FF0000000000000ULL;
            double value1_2;
            std::memcpy(&value1_2, &bits, 8);
            double value0_1 = value1_2 - 1.0;
            double res = value0_1 * scale + low;
            if (res < high) return res;
            // rounding pushed res onto `high`: shave one ulp off the scale and draw again
            uint64_t sb;
            std::memcpy(&sb, &scale, 8);
            sb -= 1;
            std::memcpy(&scale, &sb, 8);
        }
    }

    double unit() { return gen_range_f64(0.0, 1.0); }

    // rand 0.8 UniformInt<usize>::sample_single on a 64-bit target: widening multiply with a
    // rejection zone, so the number of next_u64 calls depends on the values drawn.
    uint64_t gen_range_usize(uint64_t low, uint64_t high) {
        uint64_t range = high - low;
        uint64_t zone = (range << __builtin_clzll(range)) - 1;
        for (;;) {
            uint64_t v = next_u64();
            u128 m = (u128)v * (u128)range;
            uint64_t hi = (uint64_t)(m >> 64), lo = (uint64_t)m;
            if (lo <= zone) return low + hi;
        }
    }
};

}  // namespace orc
