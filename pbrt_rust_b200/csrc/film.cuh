// Film accumulation kernel (sm_100a): deterministic per-pixel gather.
//
// Replaces Film::add_sample (src/camera/film.rs:192-249) over all samples of a frame.  One thread
// owns one film pixel and visits, in global raster order (sampler pixel row, column, sample index),
// every camera sample whose filter footprint can reach it; for each it repeats add_sample's own
// extent and table-index arithmetic and adds `filter_wt * xyz` / `filter_wt` in that order.  The
// float summation order per pixel is therefore fixed — identical from run to run, for any tile
// partition and any GPU count — and equals a sequential CPU add_sample sweep in raster order.
#pragma once
#include "scene.cuh"

struct DFilm {
  int x_start, y_start, x_count, y_count;  // film pixel extent
  float xw, yw, inv_xw, inv_yw;
  int sx0, sx1, sy0, sy1;  // sampler extent
  int spp;
};

__constant__ float c_filter_table[256];

struct FilmArgs {
  const float2* __restrict__ img;        // per sample, list order
  const float4* __restrict__ xyz;        // per sample, list order
  const int32_t* __restrict__ pix_index; // sampler-extent raster -> list position or -1
  const int32_t* __restrict__ rects;     // film pixel rects to fill
  const uint32_t* __restrict__ rect_prefix;  // prefix sums of rect areas (n_rects + 1)
  uint32_t n_rects;
  uint32_t n_pixels;  // total pixels over all rects
  float4* __restrict__ out;  // film, row-major over the film pixel extent
};

__global__ void __launch_bounds__(128)
k_film(const DFilm f, const FilmArgs a) {
  const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= a.n_pixels) return;
  // locate the rect (few rects per GPU; linear scan)
  uint32_t r = 0;
  while (r + 1 < a.n_rects && gid >= a.rect_prefix[r + 1]) ++r;
  const int rx0 = a.rects[4 * r], ry0 = a.rects[4 * r + 1], rx1 = a.rects[4 * r + 2];
  const uint32_t local = gid - a.rect_prefix[r];
  const int rw = rx1 - rx0;
  const int x = rx0 + (int)(local % (uint32_t)rw), y = ry0 + (int)(local / (uint32_t)rw);

  // Sampler pixels q whose samples (coordinates in [q, q+1]) can reach pixel (x, y).
  int qx0 = (int)ceilf((float)x - 0.5f - f.xw) - 1, qx1 = (int)floorf((float)x + 0.5f + f.xw) + 1;
  int qy0 = (int)ceilf((float)y - 0.5f - f.yw) - 1, qy1 = (int)floorf((float)y + 0.5f + f.yw) + 1;
  qx0 = max(qx0, f.sx0);
  qx1 = min(qx1, f.sx1 - 1);
  qy0 = max(qy0, f.sy0);
  qy1 = min(qy1, f.sy1 - 1);
  const int sw = f.sx1 - f.sx0;
  float X = 0.f, Y = 0.f, Z = 0.f, Wt = 0.f;
  for (int qy = qy0; qy <= qy1; ++qy) {
    // A sample of sampler pixel q has image coordinate in [q, q+1]; all of add_sample's float ops
    // are monotonic, so its pixel extent lies inside [ceil((q-0.5)-w), floor((q+0.5)+w)].
    if (y < f2i_sat(ceilf(((float)qy - 0.5f) - f.yw)) ||
        y > f2i_sat(floorf((((float)qy + 1.0f) - 0.5f) + f.yw)))
      continue;
    for (int qx = qx0; qx <= qx1; ++qx) {
      if (x < f2i_sat(ceilf(((float)qx - 0.5f) - f.xw)) ||
          x > f2i_sat(floorf((((float)qx + 1.0f) - 0.5f) + f.xw)))
        continue;
      const int32_t li = __ldg(&a.pix_index[(size_t)(qy - f.sy0) * (size_t)sw + (size_t)(qx - f.sx0)]);
      if (li < 0) continue;
      const uint64_t base = (uint64_t)li * (uint64_t)f.spp;
      for (int i = 0; i < f.spp; ++i) {
        const float2 im = __ldg(a.img + base + i);
        // film.rs:198-210
        const float dimage_x = im.x - 0.5f, dimage_y = im.y - 0.5f;
        const int x0 = max(f.x_start, f2i_sat(ceilf(dimage_x - f.xw)));
        const int x1 = min(f.x_start + f.x_count - 1, f2i_sat(floorf(dimage_x + f.xw)));
        const int y0 = max(f.y_start, f2i_sat(ceilf(dimage_y - f.yw)));
        const int y1 = min(f.y_start + f.y_count - 1, f2i_sat(floorf(dimage_y + f.yw)));
        if ((x1 - x0) < 0 || (y1 - y0) < 0) continue;
        if (x < x0 || x > x1 || y < y0 || y > y1) continue;
        // film.rs:216-224
        const float fx = ((float)x - dimage_x) * f.inv_xw * 16.0f;
        const float fy = ((float)y - dimage_y) * f.inv_yw * 16.0f;
        const int ix = min(f2i_sat(floorf(fabsf(fx))), 15);
        const int iy = min(f2i_sat(floorf(fabsf(fy))), 15);
        const float wt = c_filter_table[iy * 16 + ix];
        const float4 c = __ldg(a.xyz + base + i);
        X += wt * c.x;  // film.rs:241-244
        Y += wt * c.y;
        Z += wt * c.z;
        Wt += wt;
      }
    }
  }
  a.out[(size_t)(y - f.y_start) * (size_t)f.x_count + (size_t)(x - f.x_start)] =
      make_float4(X, Y, Z, Wt);
}
