// Host-side world-build RNG of the product: the exact stream the reference's worlds consume.
//
// The reference seeds `rand_pcg::Pcg64::seed_from_u64(seed + site)` (src/rngator.rs:27-31, src/main.rs:185)
// and draws with `Rng::gen_range` (src/worlds.rs:63,97,407,458; src/bhv.rs:127; src/textures.rs:65,142).
// Those crates (rand 0.8.3, rand_core 0.6.2, rand_pcg 0.3.0; Cargo.lock:325-366) are not vendored, so the
// arithmetic below restates their published algorithms.  Only scene construction uses this generator; the
// render streams on the device are counter-based Philox (rt_device.cuh).
#pragma once
#include <cstdint>
#include <cstring>

namespace rtb {

class WorldRng {
   public:
    explicit WorldRng(uint64_t seed) {
        // rand_core::SeedableRng::seed_from_u64 — a PCG32 generator fills the 32 seed bytes, LE words
        uint32_t words[8];
        uint64_t s = seed;
        for (int i = 0; i < 8; ++i) {
            s = s * 6364136223846793005ULL + 11634580027462260723ULL;
            uint32_t xs = (uint32_t)(((s >> 18) ^ s) >> 27);
            uint32_t rot = (uint32_t)(s >> 59);
            words[i] = (xs >> rot) | (xs << ((-rot) & 31u));
        }
        uint64_t q[4];
        for (int i = 0; i < 4; ++i) q[i] = (uint64_t)words[2 * i] | ((uint64_t)words[2 * i + 1] << 32);
        // Lcg128Xsl64::from_seed -> from_state_incr(state, incr | 1): state += incr, one step
        inc_ = (((unsigned __int128)q[3] << 64) | q[2]) | 1u;
        state_ = (((unsigned __int128)q[1] << 64) | q[0]) + inc_;
        advance();
    }

    uint64_t next_u64() {
        ++calls_;
        advance();
        // XSL-RR 128/64 output function
        uint64_t hi = (uint64_t)(state_ >> 64), lo = (uint64_t)state_;
        uint32_t rot = (uint32_t)(hi >> 58);
        uint64_t x = hi ^ lo;
        return (x >> rot) | (x << ((-rot) & 63u));
    }

    // UniformFloat<f64>::sample_single: 52 random mantissa bits -> [1,2) - 1, then value*scale + low as two
    // separately rounded operations; redraw (with scale one ulp smaller) if rounding reaches `high`.
    double range_f64(double low, double high) {
        double scale = high - low;
        while (true) {
            uint64_t bits = (next_u64() >> 12) | 0x3FF0000000000000ULL;
            double one_two;
            std::memcpy(&one_two, &bits, sizeof one_two);
            volatile double prod = (one_two - 1.0) * scale;  // volatile: keep the product rounded (no FMA)
            double r = prod + low;
            if (r < high) return r;
            uint64_t sb;
            std::memcpy(&sb, &scale, sizeof sb);
            --sb;
            std::memcpy(&scale, &sb, sizeof sb);
        }
    }
    double unit() { return range_f64(0.0, 1.0); }

    // UniformInt<usize>::sample_single (64-bit): widening multiply, reject the biased low zone
    uint64_t range_usize(uint64_t low, uint64_t high) {
        uint64_t span = high - low;
        uint64_t zone = (span << __builtin_clzll(span)) - 1;
        while (true) {
            unsigned __int128 wide = (unsigned __int128)next_u64() * span;
            if ((uint64_t)wide <= zone) return low + (uint64_t)(wide >> 64);
        }
    }

    uint64_t calls() const { return calls_; }

   private:
    void advance() {
        const unsigned __int128 mult = ((unsigned __int128)0x2360ED051FC65DA4ULL << 64) | 0x4385DF649FCCF645ULL;
        state_ = state_ * mult + inc_;
    }
    unsigned __int128 state_, inc_;
    uint64_t calls_ = 0;
};

}  // namespace rtb
