// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// pbrt_rust's RNG (src/rng.rs:6-43) is `rand::rngs::StdRng::seed_from_u64(task_idx)`.  The `rand`
// crates are NOT vendored under the reference (Cargo.lock pins rand 0.8.5, rand_chacha 0.3.1,
// rand_core 0.6.4), so their published algorithm is restated here:
//   * rand_core 0.6 `SeedableRng::seed_from_u64`: a PCG32 generator (multiplier
//     6364136223846793005, increment 11634580027462260723) fills the 32-byte seed, 4 bytes (LE)
//     per step: state = state*MUL + INC; x = ((state>>18)^state)>>27; out = rotr32(x, state>>59).
//   * rand 0.8 `StdRng` = `rand_chacha::ChaCha12Rng`: ChaCha with 12 rounds, key = seed,
//     64-bit block counter in state words 12..13 (starts at 0), 64-bit stream id in words 14..15
//     (0).  Output word w of the stream is word (w mod 16) of block (w div 16); the crate's
//     4-block buffering preserves that order, and `next_u64` = (lo = next word, hi = the one
//     after), also across refills.
//   * `gen::<f32>()` = (next_u32() >> 8) as f32 * 2^-24; `gen::<u64>()` = next_u64().
// Pins: the ChaCha core is checked against RFC 8439 §2.3.2 (20 rounds) and the all-zero-key
// ChaCha12/ChaCha20 keystream vectors; StdRng::from_seed is checked against rand 0.8's own
// `test_stdrng_construction` value.  `seed_from_u64` has no published known-answer value:
// PARITY UNPINNED at that one step.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace orc {

inline uint32_t rotl32(uint32_t v, int c) { return (v << c) | (v >> (32 - c)); }

// One ChaCha block: out = in + rounds(in).  `in` is the full 16-word state.
inline void chacha_block(const uint32_t in[16], int rounds, uint32_t out[16]) {
  uint32_t x[16];
  std::memcpy(x, in, sizeof x);
#define ORC_QR(a, b, c, d) \
  x[a] += x[b];            \
  x[d] = rotl32(x[d] ^ x[a], 16); \
  x[c] += x[d];            \
  x[b] = rotl32(x[b] ^ x[c], 12); \
  x[a] += x[b];            \
  x[d] = rotl32(x[d] ^ x[a], 8);  \
  x[c] += x[d];            \
  x[b] = rotl32(x[b] ^ x[c], 7);
  for (int i = 0; i < rounds; i += 2) {
    ORC_QR(0, 4, 8, 12)
    ORC_QR(1, 5, 9, 13)
    ORC_QR(2, 6, 10, 14)
    ORC_QR(3, 7, 11, 15)
    ORC_QR(0, 5, 10, 15)
    ORC_QR(1, 6, 11, 12)
    ORC_QR(2, 7, 8, 13)
    ORC_QR(3, 4, 9, 14)
  }
#undef ORC_QR
  for (int i = 0; i < 16; ++i) out[i] = x[i] + in[i];
}

// rand_core 0.6.4 SeedableRng::seed_from_u64 -> 8 LE u32 key words.
inline void seed_from_u64(uint64_t state, uint32_t key[8]) {
  const uint64_t MUL = 6364136223846793005ull;
  const uint64_t INC = 11634580027462260723ull;
  for (int i = 0; i < 8; ++i) {
    state = state * MUL + INC;
    uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
    uint32_t rot = (uint32_t)(state >> 59);
    key[i] = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
  }
}

// Counter-addressable ChaCha12 word stream (== StdRng's output sequence).
struct ChaChaStream {
  uint32_t key[8];
  int rounds = 12;
  uint64_t pos = 0;        // index of the next word
  uint64_t cached_block = UINT64_MAX;
  uint32_t buf[16];
  void load_block(uint64_t blk) {
    uint32_t st[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
    for (int i = 0; i < 8; ++i) st[4 + i] = key[i];
    st[12] = (uint32_t)blk;
    st[13] = (uint32_t)(blk >> 32);
    st[14] = 0;
    st[15] = 0;
    chacha_block(st, rounds, buf);
    cached_block = blk;
  }
  uint32_t word_at(uint64_t w) {
    uint64_t blk = w >> 4;
    if (blk != cached_block) load_block(blk);
    return buf[w & 15];
  }
  uint32_t next_u32() { return word_at(pos++); }
  uint64_t next_u64() {
    uint64_t lo = next_u32();
    uint64_t hi = next_u32();
    return lo | (hi << 32);
  }
};

// src/rng.rs:6-43
struct RNG {
  ChaChaStream s;
  explicit RNG(uint64_t task_idx) { seed_from_u64(task_idx, s.key); }
  static RNG from_key(const uint32_t key[8]) {
    RNG r(0);
    std::memcpy(r.s.key, key, sizeof r.s.key);
    r.s.cached_block = UINT64_MAX;
    return r;
  }
  void seek(uint64_t word) { s.pos = word; }
  uint64_t tell() const { return s.pos; }
  float random_float() { return (float)(s.next_u32() >> 8) * (1.0f / 16777216.0f); }   // :15-17
  uint64_t random_uint() { return s.next_u64() % UINT64_MAX; }                         // :19-21
  // :23-33  Fisher-Yates over `count` groups of `dims` lanes
  void shuffle(float* v, size_t len, size_t dims) {
    size_t count = len / dims;
    for (size_t i = 0; i < count; ++i) {
      size_t other = i + (size_t)(random_uint() % (uint64_t)(count - i));
      for (size_t j = 0; j < dims; ++j) std::swap(v[dims * i + j], v[dims * other + j]);
    }
  }
};

// montecarlo.rs:88-105
inline void latin_hypercube(float* samples, size_t num, size_t dim, RNG& rng) {
  float delta = 1.0f / (float)num;
  for (size_t i = 0; i < num; ++i)
    for (size_t j = 0; j < dim; ++j)
      samples[dim * i + j] = ((float)i + rng.random_float()) * delta;
  for (size_t i = 0; i < dim; ++i)
    for (size_t j = 0; j < num; ++j) {
      size_t other = j + (size_t)(rng.random_uint() % (uint64_t)(num - j));
      std::swap(samples[dim * j + i], samples[dim * other + i]);
    }
}
// montecarlo.rs:107-114
inline void stratified_sample_1d(float* samples, size_t n, RNG& rng, bool jitter) {
  float inv_tot = 1.0f / (float)n;
  for (size_t i = 0; i < n; ++i) {
    float delta = jitter ? rng.random_float() : 0.5f;
    samples[i] = ((float)i + delta) * inv_tot;
  }
}
// montecarlo.rs:116-129
inline void stratified_sample_2d(float* samples, size_t nx, size_t ny, RNG& rng, bool jitter) {
  float dx = 1.0f / (float)nx;
  float dy = 1.0f / (float)ny;
  for (size_t y = 0; y < ny; ++y)
    for (size_t x = 0; x < nx; ++x) {
      float jx = jitter ? rng.random_float() : 0.5f;
      float jy = jitter ? rng.random_float() : 0.5f;
      size_t off = 2 * (y * nx + x);
      samples[off] = ((float)x + jx) * dx;
      samples[off + 1] = ((float)y + jy) * dy;
    }
}

}  // namespace orc
