// Film::add_sample arithmetic (src/camera/film.rs:192-249) for one (sample, pixel) pair (sm_100a).
// Plain arithmetic: the same source compiles as host code (PB_HOST_CHECK, tests/devsrc/) so the CPU
// test-suite can run it against the oracle; the product runs it on the GPU.
#pragma once
#include "dmath.cuh"

struct DFilm {
  int x_start, y_start, x_count, y_count;  // film pixel extent
  float xw, yw, inv_xw, inv_yw;
  int sx0, sx1, sy0, sy1;  // sampler extent
  int spp;
};

// Does the sample at image (ix_, iy_) reach film pixel (x, y)?  If so *ti = its index into the 16 x 16
// filter table.  film.rs:198-210 (extent, clipped to the film) and :216-224 (table lookup).
PB_DEV bool film_sample_index(const DFilm& f, float ix_, float iy_, int x, int y, int* ti) {
  const float dimage_x = ix_ - 0.5f, dimage_y = iy_ - 0.5f;
  const int x0 = max(f.x_start, f2i_sat(ceilf(dimage_x - f.xw)));
  const int x1 = min(f.x_start + f.x_count - 1, f2i_sat(floorf(dimage_x + f.xw)));
  const int y0 = max(f.y_start, f2i_sat(ceilf(dimage_y - f.yw)));
  const int y1 = min(f.y_start + f.y_count - 1, f2i_sat(floorf(dimage_y + f.yw)));
  if ((x1 - x0) < 0 || (y1 - y0) < 0) return false;
  if (x < x0 || x > x1 || y < y0 || y > y1) return false;
  const float fx = ((float)x - dimage_x) * f.inv_xw * 16.0f;
  const float fy = ((float)y - dimage_y) * f.inv_yw * 16.0f;
  const int ix = min(f2i_sat(floorf(fabsf(fx))), 15);
  const int iy = min(f2i_sat(floorf(fabsf(fy))), 15);
  *ti = iy * 16 + ix;
  return true;
}
