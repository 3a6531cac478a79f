// Film accumulation kernel (sm_100a): deterministic per-pixel gather, fused with the radiance fold.
//
// Replaces Film::add_sample (src/camera/film.rs:192-249) over all samples of a frame, the light
// sum of WhittedIntegrator::li (src/integrator/whitted.rs:49-66) and Spectrum::to_xyz
// (src/spectrum.rs:37-41,443-458).  One thread owns one film pixel and visits, in global raster
// order (sampler pixel row, column, sample index), every camera sample whose filter footprint can
// reach it; for each it repeats add_sample's own extent and table-index arithmetic and adds
// `filter_wt * xyz` / `filter_wt` in that order.  The float summation order per pixel is therefore
// fixed — identical from run to run, for any tile partition and any GPU count — and equals a
// sequential CPU add_sample sweep in raster order.
//
// Neighbour skipping: raygen marks a sampler pixel in `edge` iff one of its samples has an
// add_sample extent other than exactly its own pixel.  An unmarked neighbour cannot contribute, so
// its samples are not even loaded; with the 0.5 box filter this reduces the gather to the pixel's
// own samples plus eight flag loads.  Wide filters mark every pixel and take the full gather.
#pragma once
#include "scene.cuh"

#define PB_MAX_FOLD_LIGHTS 64
struct DFold {  // how the per-sample radiance terms fold into L (light order)
  uint32_t rad_slots, le_slot, n_lights;
  uint16_t ns[PB_MAX_FOLD_LIGHTS];   // samples of light i (1 for point/spot)
  uint8_t area[PB_MAX_FOLD_LIGHTS];  // 1 = area light (averaged over its samples)
};

#include "film_math.cuh"  // DFilm, film_sample_index

#ifdef PB_HOST_CHECK
static float c_filter_table[256];  // host check: a plain array the harness fills
#else
__constant__ float c_filter_table[256];
#endif

struct FilmArgs {
  const float2* __restrict__ img;         // per sample, list order
  const float4* __restrict__ rad;         // per sample x rad_slots radiance terms (rgb)
  const uint32_t* __restrict__ offsets;   // HaltonSampler: list pixel li owns samples [offsets[li], offsets[li+1]); else NULL (li * spp)
  const uint32_t* __restrict__ edge;      // per list pixel
  const int32_t* __restrict__ pix_index;  // sampler-extent raster -> list position or -1
  const int32_t* __restrict__ rects;      // film pixel rects to fill
  const uint32_t* __restrict__ rect_prefix;  // prefix sums of rect areas (n_rects + 1)
  uint32_t n_rects;
  uint32_t n_pixels;  // total pixels over all rects
  uint32_t first, count;  // this launch covers pixel ordinals [first, first + count)
  float4* __restrict__ out;  // film, row-major over the film pixel extent
  uint32_t* nan_count;
};

// L = Le + sum over lights; to_xyz.  Returns true if L has a NaN (sampler_renderer.rs:105 intent).
PB_DEV bool fold_radiance(const DFold& fd, const float4* __restrict__ r, float* X, float* Y,
                          float* Z) {
  f3 L = mk3(0.f, 0.f, 0.f);
  uint32_t slot = 0;
  if (fd.le_slot) {
    const float4 le = __ldg(r);
    L = mk3(le.x, le.y, le.z);
    slot = 1;
  }
  for (uint32_t li = 0; li < fd.n_lights; ++li) {
    if (fd.area[li]) {
      const uint32_t ns = fd.ns[li];
      f3 Ld = mk3(0.f, 0.f, 0.f);
      for (uint32_t s = 0; s < ns; ++s) {
        const float4 c = __ldg(r + slot++);
        Ld = Ld + mk3(c.x, c.y, c.z);
      }
      const float fns = (float)ns;
      L = L + mk3(Ld.x / fns, Ld.y / fns, Ld.z / fns);
    } else {
      const float4 c = __ldg(r + slot++);
      L = L + mk3(c.x, c.y, c.z);
    }
  }
  *X = 0.412453f * L.x + 0.357580f * L.y + 0.180423f * L.z;
  *Y = 0.212671f * L.x + 0.715160f * L.y + 0.072169f * L.z;
  *Z = 0.019334f * L.x + 0.119193f * L.y + 0.950227f * L.z;
  return isnan(L.x) || isnan(L.y) || isnan(L.z);
}

// One film pixel (ordinal `gid` over the rects of this call): the whole gather.  Written as a
// function of the ordinal so that the host check (tests/devsrc/) can run it pixel by pixel.
PB_DEV void film_pixel(const DFilm& f, const DFold& fd, const FilmArgs& a, uint32_t gid) {
  // locate the rect (few rects per GPU; binary search over the prefix sums)
  uint32_t lo = 0, hi = a.n_rects;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (gid >= a.rect_prefix[mid]) lo = mid; else hi = mid;
  }
  const uint32_t r = lo;
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
  uint32_t nans = 0;
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
      const bool own = (qx == x) && (qy == y);
      if (!own && __ldg(&a.edge[li]) == 0u) continue;  // cannot reach any pixel but its own
      uint64_t base = (uint64_t)li * (uint64_t)f.spp;
      int cnt = f.spp;
      if (a.offsets) {  // variable samples per pixel (halton.cuh)
        base = __ldg(&a.offsets[li]);
        cnt = (int)(__ldg(&a.offsets[li + 1]) - (uint32_t)base);
      }
      for (int i = 0; i < cnt; ++i) {
        const float2 im = __ldg(a.img + base + i);
        int ti;
        if (!film_sample_index(f, im.x, im.y, x, y, &ti)) continue;
        const float wt = c_filter_table[ti];
        float cx, cy, cz;
        const bool bad = fold_radiance(fd, a.rad + (base + i) * fd.rad_slots, &cx, &cy, &cz);
        if (bad && own) ++nans;
        X += wt * cx;  // film.rs:241-244
        Y += wt * cy;
        Z += wt * cz;
        Wt += wt;
      }
    }
  }
  if (nans) atomicAdd(a.nan_count, nans);
  a.out[(size_t)(y - f.y_start) * (size_t)f.x_count + (size_t)(x - f.x_start)] =
      make_float4(X, Y, Z, Wt);
}

#ifndef PB_HOST_CHECK
__global__ void __launch_bounds__(128)
k_film(const DFilm f, const DFold fd, const FilmArgs a) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= a.count) return;
  film_pixel(f, fd, a, a.first + tid);
}

// Film::write_image pixel pipeline (film.rs:331-340 as intended, write_img film.rs:21-23).
__global__ void __launch_bounds__(256)
k_film_develop(const float4* __restrict__ film, uint64_t n, float* __restrict__ out_rgb,
               uint8_t* __restrict__ out_rgb8) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = __ldg(film + i);
  float r = 3.240479f * p.x - 1.37150f * p.y - 0.498535f * p.z;  // spectrum.rs:31-35
  float g = -0.969256f * p.x + 1.875991f * p.y + 0.041556f * p.z;
  float b = 0.055648f * p.x - 0.204043f * p.y + 1.057311f * p.z;
  if (p.w != 0.0f) {
    const float inv = 1.0f / p.w;
    r = fmaxf(r * inv, 0.0f);
    g = fmaxf(g * inv, 0.0f);
    b = fmaxf(b * inv, 0.0f);
  }
  if (out_rgb) {
    out_rgb[3 * i] = r;
    out_rgb[3 * i + 1] = g;
    out_rgb[3 * i + 2] = b;
  }
  if (out_rgb8) {
    auto to_byte = [](float v) -> uint8_t {
      float q = 255.0f * powf(v, 1.0f / 2.2f) + 0.5f;
      q = q < 0.0f ? 0.0f : (q > 255.0f ? 255.0f : q);
      return isnan(q) ? (uint8_t)0 : (uint8_t)q;
    };
    out_rgb8[3 * i] = to_byte(r);
    out_rgb8[3 * i + 1] = to_byte(g);
    out_rgb8[3 * i + 2] = to_byte(b);
  }
}
#endif  // PB_HOST_CHECK
