// HaltonSampler arithmetic (sm_100a): radical_inverse (src/montecarlo.rs:7-20, f64 as written) and
// the candidate -> image position map of HaltonSampler::get_more_samples (src/sampler/halton.rs:57-76).
// Plain arithmetic on plain structs: the same source compiles as host code (PB_HOST_CHECK,
// tests/devsrc/) so the CPU test-suite can run it against the oracle; the product runs it on the GPU.
#pragma once
#include "dmath.cuh"

#ifdef PB_HOST_CHECK
#define PB_HTABLE static const
#else
#define PB_HTABLE static __device__ const
#endif

struct DHaltonTask {
  int x0, x1, y0, y1;            // the task's sampler sub-window (sampler/base.rs:29-48)
  float delta;                   // lerp_delta = dy.max(dx) (halton.rs:62-66)
  uint32_t pad;
  unsigned long long first;      // global ordinal of this task's candidate 0
  unsigned long long wanted;     // max(dx, dy)^2 * samples_per_pixel (halton.rs:20-27)
};

#define PB_HALTON_MAX_LIGHT_PAIRS 16
PB_HTABLE unsigned int pb_halton_primes[40] = {
    2,  3,  5,  7,  11, 13, 17, 19, 23, 29,  31,  37,  41,  43,  47,  53,  59,  61,  67,  71,
    73, 79, 83, 89, 97, 101, 103, 107, 109, 113, 127, 131, 137, 139, 149, 151, 157, 163, 167, 173};

// montecarlo.rs:7-20
PB_DEV double radical_inverse_(unsigned long long n, unsigned int b) {
  double v = 0.0;
  const double inv_base = 1.0 / (double)b;
  double aib = 1.0;
  while (n > 0) {
    const double d = (double)(n % b);
    n /= b;
    aib *= inv_base;
    v += d * aib;
  }
  return v;
}

// halton.rs:57-76: image position of candidate i; false = skipped (outside the window)
PB_DEV bool halton_image(const DHaltonTask& t, unsigned long long i, float* ix, float* iy) {
  const float u = (float)radical_inverse_(i, 3u);
  const float v = (float)radical_inverse_(i, 2u);
  const float xs = (float)t.x0, ys = (float)t.y0;
  const float image_x = lerpf_(xs, xs + t.delta, u);
  const float image_y = lerpf_(ys, ys + t.delta, v);
  if (image_x >= (float)t.x1 || image_y >= (float)t.y1) return false;
  *ix = image_x;
  *iy = image_y;
  return true;
}

