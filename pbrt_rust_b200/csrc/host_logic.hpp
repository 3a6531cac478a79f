// Host-side logic the back end needs inside the library (plain C++, no CUDA): the per-task sampler
// sub-windows and RNG keys that SamplerRenderer derives before rendering, and the pixel work list.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace pbh {

// Rust `f32 as i32` / `as usize` (saturating; NaN -> 0)
inline int32_t sat_i32(float x) {
  if (x != x) return 0;
  if (x >= 2147483648.0f) return INT32_MAX;
  if (x <= -2147483648.0f) return INT32_MIN;
  return (int32_t)x;
}
inline uint64_t sat_usize(float x) {
  if (x != x || x <= 0.0f) return 0;
  if (x >= 18446744073709551616.0f) return UINT64_MAX;
  return (uint64_t)x;
}

// rand_core 0.6 SeedableRng::seed_from_u64 (PCG32 expansion) — the key of
// StdRng::seed_from_u64(task_idx), src/rng.rs:11-13.
inline void task_key(uint64_t task_idx, uint32_t key[8]) {
  uint64_t st = task_idx;
  for (int i = 0; i < 8; ++i) {
    st = st * 6364136223846793005ull + 11634580027462260723ull;
    const uint32_t xs = (uint32_t)(((st >> 18) ^ st) >> 27);
    const uint32_t rot = (uint32_t)(st >> 59);
    key[i] = (xs >> rot) | (xs << ((32u - rot) & 31u));
  }
}

// src/utils/mod.rs:171-205
inline void crop_window(uint64_t num, uint64_t count, float aspect, float w[4]) {
  auto split = [](uint64_t cnt, uint64_t asp, uint64_t* nx, uint64_t* ny) {
    uint64_t x = 1, y = cnt;
    while ((y % 2) == 0 && 2 * asp * x < y) {
      y /= 2;
      x *= 2;
    }
    *nx = x;
    *ny = y;
  };
  uint64_t nx, ny;
  if (aspect < 1.0f) {
    split(count, sat_usize(1.0f / aspect), &nx, &ny);
  } else {
    uint64_t a, b;
    split(count, sat_usize(aspect), &a, &b);
    nx = b;
    ny = a;
  }
  const uint64_t xo = num % nx, yo = num / nx;
  w[0] = (float)xo / (float)nx;
  w[1] = (float)(xo + 1) / (float)nx;
  w[2] = (float)yo / (float)ny;
  w[3] = (float)(yo + 1) / (float)ny;
}

// src/sampler/base.rs:29-48
inline void sampler_sub_window(const int32_t ext[4], uint64_t num, uint64_t count, int32_t out[4]) {
  const uint64_t dx = (uint64_t)(ext[1] - ext[0]), dy = (uint64_t)(ext[3] - ext[2]);
  const float aspect = (float)dx / (float)dy;
  float t[4];
  crop_window(num, count, aspect, t);
  const float psx = (float)ext[0], pex = (float)ext[1], psy = (float)ext[2], pey = (float)ext[3];
  auto lerp = [](float a, float b, float u) { return a * (1.0f - u) + b * u; };
  out[0] = sat_i32(lerp(psx, pex, t[0]));
  out[1] = sat_i32(lerp(psx, pex, t[1]));
  out[2] = sat_i32(lerp(psy, pey, t[2]));
  out[3] = sat_i32(lerp(psy, pey, t[3]));
}

}  // namespace pbh
