// Film accumulation kernel (sm_100a): deterministic per-pixel gather over per-sample radiance records.
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
struct DFold {  // how the per-sample radiance terms of a multi-light scene fold into L (light order)
  uint32_t slots, n_lights;
  uint16_t ns[PB_MAX_FOLD_LIGHTS];   // samples of light i (1 for point/spot)
  uint8_t area[PB_MAX_FOLD_LIGHTS];  // 1 = area light (averaged over its samples)
};

#include "film_math.cuh"  // DFilm, film_sample_index

// Per camera sample the frame keeps ONE float4 "radiance record" (list order; a ring over list pixels
// when the frame is larger than the wavefront budget, DESIGN.md "Radiance path"):
//   xyz = v, w = bits(e): L = (e ? Le(light e - 1) : 0) + v
// With a single light slot v is that slot's term f * Li * |wi.n| / pdf, written by k_shade and
// zeroed by the any-hit kernel when the shadow ray is occluded, so L = Le + v is exactly
// whitted.rs:49-66 for one light.  With several slots k_fold has already folded everything
// (Le included, in light order) and e = 0.
struct FilmArgs {
  const float2* __restrict__ img;         // per sample, list order (ring-addressed like `rec`)
  const float4* __restrict__ rec;         // radiance records; NULL: the scene has no light (L = 0)
  const pbrtb200_light* __restrict__ lights;  // Le lookup for e != 0
  const uint32_t* __restrict__ offsets;   // HaltonSampler: list pixel li owns samples [offsets[li], offsets[li+1]); else NULL (li * spp)
  const uint32_t* __restrict__ edge;      // per list pixel
  const int32_t* __restrict__ pix_index;  // sampler-extent raster -> list position or -1
  const int32_t* __restrict__ rects;      // film pixel rects to fill
  const uint32_t* __restrict__ rect_prefix;  // prefix sums of rect areas (n_rects + 1)
  uint32_t n_rects;
  uint32_t n_pixels;  // total pixels over all rects
  uint32_t first, count;  // this launch covers pixel ordinals [first, first + count)
  uint32_t pixel_mask;    // ring over list pixels: list pixel li lives at slot li & pixel_mask (all ones: no ring)
  float4* __restrict__ out;  // film, row-major over the film pixel extent
  float table[256];          // the film's 16 x 16 filter table (film.rs:99-110), per launch: two
                             // contexts with different filters may render concurrently
};

// XYZ of a radiance record (spectrum.rs:37-41 for RGB spectra).
PB_DEV void record_xyz(const FilmArgs& a, float4 r, float* X, float* Y, float* Z) {
  f3 L = mk3(r.x, r.y, r.z);
  const uint32_t e = __float_as_uint(r.w);
  if (e) {  // the sample hit an emitter: L = Le + v, Le first as in whitted.rs:46-66
    const pbrtb200_light* lt = a.lights + (e - 1u);
    L = mk3(__ldg(&lt->intensity[0]), __ldg(&lt->intensity[1]), __ldg(&lt->intensity[2])) + L;
  }
  *X = 0.412453f * L.x + 0.357580f * L.y + 0.180423f * L.z;
  *Y = 0.212671f * L.x + 0.715160f * L.y + 0.072169f * L.z;
  *Z = 0.019334f * L.x + 0.119193f * L.y + 0.950227f * L.z;
}

// L = Le + sum over lights (multi-slot scenes), whitted.rs:46-66 with D10's per-light average.
// terms: `slots` float4 per sample, term j = (c_j, j == 0 ? bits(e) : 0).  Returns true if L has a NaN
// (sampler_renderer.rs:105 intent).
PB_DEV bool fold_terms(const DFold& fd, const pbrtb200_light* __restrict__ lights, const float4* __restrict__ r, float4* out) {
  const float4 first = ld_stream(r);
  f3 L = mk3(0.f, 0.f, 0.f);
  const uint32_t e = __float_as_uint(first.w);
  if (e) L = mk3(lights[e - 1u].intensity[0], lights[e - 1u].intensity[1], lights[e - 1u].intensity[2]);
  uint32_t slot = 0;
  for (uint32_t li = 0; li < fd.n_lights; ++li) {
    if (fd.area[li]) {
      const uint32_t ns = fd.ns[li];
      f3 Ld = mk3(0.f, 0.f, 0.f);
      for (uint32_t s = 0; s < ns; ++s) {
        const float4 c = ld_stream(r + slot++);
        Ld = Ld + mk3(c.x, c.y, c.z);
      }
      const float fns = (float)ns;
      L = L + mk3(Ld.x / fns, Ld.y / fns, Ld.z / fns);
    } else {
      const float4 c = ld_stream(r + slot++);
      L = L + mk3(c.x, c.y, c.z);
    }
  }
  *out = make_float4(L.x, L.y, L.z, 0.f);
  return isnan(L.x) || isnan(L.y) || isnan(L.z);
}

// One film pixel (ordinal `gid` over the rects of this call): the whole gather.  Written as a
// function of the ordinal so that the host check (tests/devsrc/) can run it pixel by pixel.
PB_DEV void film_pixel(const DFilm& f, const FilmArgs& a, uint32_t gid) {
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
      uint64_t base = (uint64_t)((uint32_t)li & a.pixel_mask) * (uint64_t)f.spp;
      int cnt = f.spp;
      if (a.offsets) {  // variable samples per pixel (halton.cuh)
        base = __ldg(&a.offsets[li]);
        cnt = (int)(__ldg(&a.offsets[li + 1]) - (uint32_t)base);
      }
      for (int i = 0; i < cnt; ++i) {
        const float2 im = __ldg(a.img + base + i);
        int ti;
        if (!film_sample_index(f, im.x, im.y, x, y, &ti)) continue;
        const float wt = a.table[ti];
        float cx = 0.f, cy = 0.f, cz = 0.f;
        if (a.rec) record_xyz(a, __ldg(a.rec + base + i), &cx, &cy, &cz);
        X += wt * cx;  // film.rs:241-244
        Y += wt * cy;
        Z += wt * cz;
        Wt += wt;
      }
    }
  }
  st_stream(a.out + ((size_t)(y - f.y_start) * (size_t)f.x_count + (size_t)(x - f.x_start)), make_float4(X, Y, Z, Wt));
}

#ifndef PB_HOST_CHECK
__global__ void __launch_bounds__(128)
k_film(const DFilm f, const FilmArgs a) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= a.count) return;
  film_pixel(f, a, a.first + tid);
}

// Multi-slot scenes: folds the `slots` radiance terms of each sample of a chunk into its radiance
// record (one thread per sample) and counts NaN radiances once per sample.
struct FoldArgs {
  const float4* __restrict__ terms;  // chunk-local: slots per sample
  const pbrtb200_light* __restrict__ lights;
  float4* __restrict__ rec;          // the chunk's slice of the record buffer
  uint64_t n;
  uint32_t* nan_count;
};
__global__ void __launch_bounds__(256)
k_fold(const DFold fd, const FoldArgs a) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  float4 r;
  if (fold_terms(fd, a.lights, a.terms + i * fd.slots, &r)) atomicAdd(a.nan_count, 1u);
  a.rec[i] = r;
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
