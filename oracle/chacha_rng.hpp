// chacha_rng.hpp — TEST INFRASTRUCTURE (oracle), not product code.
//
// Restatement of `rand::rngs::StdRng` as pinned by the reference's Cargo.lock
// (rand 0.8.4 -> rand_chacha 0.3.1 -> ChaCha12, rand_core 0.6.3; Cargo.lock:315-338).
// Those crates are NOT vendored under /root/reference, so this file restates their
// published algorithms:
//   * ChaCha (D. J. Bernstein) with 12 rounds, 64-bit block counter in state words
//     12-13, stream id 0 in words 14-15, four blocks generated per refill
//     (rand_chacha `ChaCha12Core`, BlockRng with a 64-word buffer);
//   * `SeedableRng::seed_from_u64` (rand_core): PCG32 expansion of the u64 into the
//     32-byte key;
//   * the `Standard`/`Uniform` distributions used at the reference call sites
//     (main.rs:964,968-969; math.rs:9-11,32,40-41,56-57; camera.rs:71; pdf.rs:63;
//      hittable.rs:153; aarect.rs:142-144; bvh.rs:84; material.rs:146;
//      constant_medium.rs:85).
// PARITY STATUS: the ChaCha core is pinned by the RFC 7539 §2.3.2 block vector and the
// all-zero-key ChaCha20 keystream (tests/test_oracle_rng.py); the rand-specific glue
// (seed_from_u64, buffer order, float conversions) is "parity unpinned" — the reference
// holds no RNG test vector and the renderer parity is statistical by construction.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>

namespace oracle {

inline uint32_t rotl32(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }

// One ChaCha block with `rounds` rounds. state[0..3] are the constants.
inline void chacha_block(const uint32_t in[16], int rounds, uint32_t out[16]) {
    uint32_t x[16];
    std::memcpy(x, in, sizeof(x));
#define ORACLE_QR(a, b, c, d)                                                                                          \
    x[a] += x[b], x[d] = rotl32(x[d] ^ x[a], 16);                                                                      \
    x[c] += x[d], x[b] = rotl32(x[b] ^ x[c], 12);                                                                      \
    x[a] += x[b], x[d] = rotl32(x[d] ^ x[a], 8);                                                                       \
    x[c] += x[d], x[b] = rotl32(x[b] ^ x[c], 7);
    for (int r = 0; r < rounds; r += 2) {
        ORACLE_QR(0, 4, 8, 12) ORACLE_QR(1, 5, 9, 13) ORACLE_QR(2, 6, 10, 14) ORACLE_QR(3, 7, 11, 15)
        ORACLE_QR(0, 5, 10, 15) ORACLE_QR(1, 6, 11, 12) ORACLE_QR(2, 7, 8, 13) ORACLE_QR(3, 4, 9, 14)
    }
#undef ORACLE_QR
    for (int i = 0; i < 16; ++i) out[i] = x[i] + in[i];
}

class StdRng {
  public:
    static constexpr int ROUNDS = 12;
    static constexpr int BUF_WORDS = 64; // four blocks per refill

    StdRng() { std::memset(key_, 0, sizeof(key_)); }

    static StdRng from_seed(const uint8_t seed[32]) {
        StdRng r;
        for (int i = 0; i < 8; ++i)
            r.key_[i] = uint32_t(seed[4 * i]) | (uint32_t(seed[4 * i + 1]) << 8) | (uint32_t(seed[4 * i + 2]) << 16) |
                        (uint32_t(seed[4 * i + 3]) << 24);
        r.counter_ = 0;
        r.index_ = BUF_WORDS;
        return r;
    }

    static StdRng seed_from_u64(uint64_t state) {
        const uint64_t MUL = 6364136223846793005ull, INC = 11634580027462260723ull;
        uint8_t seed[32];
        for (int c = 0; c < 8; ++c) {
            state = state * MUL + INC;
            uint32_t xorshifted = uint32_t(((state >> 18) ^ state) >> 27);
            uint32_t rot = uint32_t(state >> 59);
            uint32_t x = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
            seed[4 * c] = uint8_t(x), seed[4 * c + 1] = uint8_t(x >> 8), seed[4 * c + 2] = uint8_t(x >> 16), seed[4 * c + 3] = uint8_t(x >> 24);
        }
        return from_seed(seed);
    }

    uint32_t next_u32() {
        if (index_ >= BUF_WORDS) refill();
        return buf_[index_++];
    }
    uint64_t next_u64() {
        if (index_ < BUF_WORDS - 1) {
            uint64_t lo = buf_[index_], hi = buf_[index_ + 1];
            index_ += 2;
            return (hi << 32) | lo;
        }
        if (index_ >= BUF_WORDS) {
            refill();
            uint64_t lo = buf_[0], hi = buf_[1];
            index_ = 2;
            return (hi << 32) | lo;
        }
        uint64_t lo = buf_[BUF_WORDS - 1];
        refill();
        uint64_t hi = buf_[0];
        index_ = 1;
        return (hi << 32) | lo;
    }

    // rng.gen::<f64>()  — Standard: 53 random bits scaled into [0,1)
    double gen_f64() { return double(next_u64() >> 11) * (1.0 / 9007199254740992.0); }
    // rng.gen::<bool>() — Standard: sign bit of a u32
    bool gen_bool() { return int32_t(next_u32()) < 0; }
    // rng.gen_range(low..high) for f64 — UniformFloat::sample_single
    double gen_range(double low, double high) {
        double scale = high - low;
        for (;;) {
            uint64_t bits = (next_u64() >> 12) | 0x3FF0000000000000ull;
            double value1_2;
            std::memcpy(&value1_2, &bits, sizeof(double));
            double res = (value1_2 - 1.0) * scale + low;
            if (res < high) return res;
            scale = std::nextafter(scale, 0.0);
        }
    }
    // rng.gen_range(0..=high) for usize — UniformInt::sample_single_inclusive (64-bit lanes)
    uint64_t gen_range_inclusive_u64(uint64_t high) {
        uint64_t range = high + 1;
        if (range == 0) return next_u64();
        uint64_t zone = (range << __builtin_clzll(range)) - 1;
        for (;;) {
            unsigned __int128 m = (unsigned __int128)next_u64() * range;
            if (uint64_t(m) <= zone) return uint64_t(m >> 64);
        }
    }
    // gen_index(rng, ubound) of rand::seq — u32 lanes when the bound fits
    uint32_t gen_index(uint32_t ubound) {
        uint32_t range = ubound;
        uint32_t zone = (range << __builtin_clz(range)) - 1;
        for (;;) {
            uint64_t m = uint64_t(next_u32()) * range;
            if (uint32_t(m) <= zone) return uint32_t(m >> 32);
        }
    }

  private:
    void refill() {
        uint32_t st[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
        for (int i = 0; i < 8; ++i) st[4 + i] = key_[i];
        st[14] = 0, st[15] = 0;
        for (int b = 0; b < 4; ++b) {
            uint64_t c = counter_ + uint64_t(b);
            st[12] = uint32_t(c), st[13] = uint32_t(c >> 32);
            chacha_block(st, ROUNDS, buf_ + 16 * b);
        }
        counter_ += 4;
        index_ = 0;
    }
    uint32_t key_[8];
    uint64_t counter_ = 0;
    uint32_t buf_[BUF_WORDS];
    int index_ = BUF_WORDS;
};

// Philox4x32-10 (Salmon et al. 2011) — used ONLY to replay the device's keyed
// free-flight draws in the closest-hit parity check (rt1w_trace_closest).
inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = uint64_t(0xD2511F53u) * c0, p1 = uint64_t(0xCD9E8D57u) * c2;
        uint32_t n0 = uint32_t(p1 >> 32) ^ c1 ^ k0, n1 = uint32_t(p1), n2 = uint32_t(p0 >> 32) ^ c3 ^ k1, n3 = uint32_t(p0);
        c0 = n0, c1 = n1, c2 = n2, c3 = n3;
        k0 += 0x9E3779B9u, k1 += 0xBB67AE85u;
    }
    out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
}

} // namespace oracle
