// HaltonSampler on the device (sm_100a).
//
// Replaces HaltonSampler::{new, get_sub_sampler, get_more_samples} (src/sampler/halton.rs:17-108)
// with radical_inverse (src/montecarlo.rs:7-20, f64 arithmetic as written).  The sampler's camera
// samples are a pure function of (task window, candidate index): candidate i of a task maps to
// image (u, v) = (ri(i, 3), ri(i, 2)) stretched over the window's bounding square and is skipped
// when it falls outside the window; lens and time use the index AFTER the increment (:72-87).
// So every candidate is evaluated independently, one thread each.
//
// What a Halton frame lacks is a fixed number of samples per pixel.  The per-sample wavefront
// buffers are therefore indexed through per-pixel offsets (an exclusive scan of the counts, done
// once per pixel list on the host): list pixel li owns samples [offsets[li], offsets[li + 1]), in
// candidate-index order (one task owns a pixel, so that is the reference's generation order
// restricted to the pixel).  Three passes:
//   k_halton_bin<0>  count the accepted candidates per home pixel              (atomics: counts only)
//   k_halton_bin<1>  scatter the candidate indices into the pixel's range      (arbitrary order)
//   k_halton_samples per pixel: sort its indices, evaluate the camera samples  (deterministic order)
// Light-sample floats are oracle-defined for this sampler: pair q of a camera sample
// is (ri(i + 1, prime[5 + 2q]), ri(i + 1, prime[6 + 2q])) — as written the reference's 1D / 2D
// sample arrays panic (halton.rs:96-107 hands latin_hypercube the slice before the offset).
#pragma once
#include "scene.cuh"
#include "trace.cuh"  // camera_ray, DCamera
#include "halton_math.cuh"  // DHaltonTask, radical_inverse_, halton_image

struct HaltonArgs {
  const DHaltonTask* __restrict__ tasks;
  uint32_t n_tasks;
  unsigned long long n_candidates;
  const int32_t* __restrict__ pix_index;  // sampler-extent raster -> list position or -1
  int sx0, sy0, sw;
  uint32_t* counts;   // per list pixel: accepted candidates (pass 0)
  uint32_t* fill;     // per list pixel: slots handed out (pass 1)
  const uint32_t* __restrict__ offsets;  // per list pixel (+1): first sample of the pixel (pass 1)
  uint32_t* idx;      // per sample: candidate index within its task
};

template <int PASS>
__global__ void __launch_bounds__(256) k_halton_bin(const HaltonArgs a) {
  const unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= a.n_candidates) return;
  uint32_t lo = 0, hi = a.n_tasks;  // last task whose first candidate is <= g
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (g >= a.tasks[mid].first) lo = mid; else hi = mid;
  }
  const DHaltonTask t = a.tasks[lo];
  const unsigned long long i = g - t.first;
  if (i >= t.wanted) return;  // (empty tasks share their `first` with the next one)
  float ix, iy;
  if (!halton_image(t, i, &ix, &iy)) return;
  // home pixel, clamped into the window: binning only, Film::add_sample decides the coverage
  int px = f2i_sat(floorf(ix)), py = f2i_sat(floorf(iy));
  px = px < t.x0 ? t.x0 : (px > t.x1 - 1 ? t.x1 - 1 : px);
  py = py < t.y0 ? t.y0 : (py > t.y1 - 1 ? t.y1 - 1 : py);
  const int32_t li = __ldg(&a.pix_index[(size_t)(py - a.sy0) * (size_t)a.sw + (size_t)(px - a.sx0)]);
  if (li < 0) return;  // not needed by the tiles of this call
  if (PASS == 0) {
    atomicAdd(&a.counts[li], 1u);
  } else {
    const uint32_t s = atomicAdd(&a.fill[li], 1u);
    const uint32_t o0 = __ldg(&a.offsets[li]), o1 = __ldg(&a.offsets[li + 1]);
    if (s < o1 - o0) a.idx[(size_t)o0 + s] = (uint32_t)i;
  }
}

struct HaltonSampleArgs {
  const DHaltonTask* __restrict__ tasks;
  const DPixel* __restrict__ pixels;
  uint64_t n_pixels;
  const uint32_t* __restrict__ offsets;  // n_pixels + 1
  uint32_t* idx;
  float2* __restrict__ img;
  float2* __restrict__ lens;    // may be NULL
  float* __restrict__ time;     // may be NULL
  float2* __restrict__ lightu;  // may be NULL; light_pairs float2 per slot
  uint32_t light_pairs;
  // per list pixel, may be NULL (primary_hits): != 0 iff some sample of the pixel has an add_sample
  // extent (film.rs:198-210) other than exactly the pixel itself — the film gather's neighbour skip
  uint32_t* __restrict__ edge;
  float xw, yw;  // filter half-widths for the edge flags
  float sopen, sclose;
};

__global__ void __launch_bounds__(128) k_halton_samples(const HaltonSampleArgs a) {
  const uint64_t li = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= a.n_pixels) return;
  const uint64_t base = a.offsets[li];
  const uint32_t cnt = a.offsets[li + 1] - a.offsets[li];
  uint32_t* q = a.idx + base;
  for (uint32_t s = 1; s < cnt; ++s) {  // insertion sort: candidate-index order
    const uint32_t v = q[s];
    uint32_t k = s;
    while (k > 0 && q[k - 1] > v) {
      q[k] = q[k - 1];
      --k;
    }
    q[k] = v;
  }
  const DPixel px = a.pixels[li];
  const DHaltonTask t = a.tasks[px.task & PB_PIXEL_TASK_MASK];
  bool edge = false;
  for (uint32_t s = 0; s < cnt; ++s) {
    const unsigned long long i = q[s];
    float2 im;
    halton_image(t, i, &im.x, &im.y);
    if (a.edge) {
      const float dimage_x = im.x - 0.5f, dimage_y = im.y - 0.5f;
      const int x0 = f2i_sat(ceilf(dimage_x - a.xw)), x1 = f2i_sat(floorf(dimage_x + a.xw));
      const int y0 = f2i_sat(ceilf(dimage_y - a.yw)), y1 = f2i_sat(floorf(dimage_y + a.yw));
      edge |= !(x0 == px_x(px) && x1 == px_x(px) && y0 == px_y(px) && y1 == px_y(px));
    }
    const unsigned long long cur = i + 1ull;  // halton.rs:72: the increment precedes the lens / time dimensions
    a.img[base + s] = im;
    if (a.lens) a.lens[base + s] = make_float2((float)radical_inverse_(cur, 5u), (float)radical_inverse_(cur, 7u));
    if (a.time) a.time[base + s] = lerpf_(a.sopen, a.sclose, (float)radical_inverse_(cur, 11u));
    if (a.lightu)
      for (uint32_t p = 0; p < a.light_pairs; ++p)
        a.lightu[(base + s) * a.light_pairs + p] =
            make_float2((float)radical_inverse_(cur, pb_halton_primes[5 + 2 * p]),
                        (float)radical_inverse_(cur, pb_halton_primes[6 + 2 * p]));
  }
  if (a.edge) a.edge[li] = edge ? 1u : 0u;
}

// primary_hits outputs of a Halton frame: compact per-sample buffers -> the padded raster layout
// [((y - y0) * w + (x - x0)) * cap + slot]; unused slots get prim = MISS and NaN image coordinates.
__global__ void k_scatter_halton(const DPixel* __restrict__ pixels, uint64_t n_pixels, uint32_t cap,
                                 const uint32_t* __restrict__ offsets, int x0, int y0, int w,
                                 const float2* __restrict__ img, const float2* __restrict__ lens,
                                 const float* __restrict__ time, const pbrtb200_hit16* __restrict__ hits,
                                 const DCamera cam, pbrtb200_hit16* __restrict__ out_hits,
                                 float* __restrict__ out_samples, pbrtb200_ray32* __restrict__ out_rays) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_pixels * cap) return;
  const uint64_t li = g / cap;
  const uint32_t slot = (uint32_t)(g - li * cap);
  const DPixel px = pixels[li];
  const uint64_t o = ((uint64_t)(px_y(px) - y0) * (uint64_t)w + (uint64_t)(px_x(px) - x0)) * cap + slot;
  const uint32_t cnt = offsets[li + 1] - offsets[li];
  const bool real = slot < cnt;
  const uint64_t s = (uint64_t)offsets[li] + slot;
  const float qnan = __int_as_float(0x7fc00000);
  if (out_hits) {
    pbrtb200_hit16 h;
    h.prim = PBRTB200_MISS;
    h.t = h.b1 = h.b2 = 0.f;
    out_hits[o] = real ? hits[s] : h;
  }
  const float2 im = real ? img[s] : make_float2(qnan, qnan);
  const float2 ln = (real && lens) ? lens[s] : make_float2(0.f, 0.f);
  if (out_samples) {
    float* q = out_samples + 5 * o;
    q[0] = im.x;
    q[1] = im.y;
    q[2] = ln.x;
    q[3] = ln.y;
    q[4] = (real && time) ? time[s] : 0.f;
  }
  if (out_rays) {
    pbrtb200_ray32 r;
    r.o[0] = r.o[1] = r.o[2] = qnan; r.mint = 0.f;
    r.d[0] = r.d[1] = r.d[2] = qnan; r.maxt = PB_F32_MAX;
    if (real) {
      f3 ro, rd;
      camera_ray(cam, im.x, im.y, ln.x, ln.y, &ro, &rd, nullptr);
      r.o[0] = ro.x; r.o[1] = ro.y; r.o[2] = ro.z;
      r.d[0] = rd.x; r.d[1] = rd.y; r.d[2] = rd.z;
    }
    out_rays[o] = r;
  }
}
