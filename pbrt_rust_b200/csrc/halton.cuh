// HaltonSampler on the device (sm_100a).
//
// Replaces HaltonSampler::{new, get_sub_sampler, get_more_samples} (src/sampler/halton.rs:17-108)
// with radical_inverse (src/montecarlo.rs:7-20, f64 arithmetic as written).  The sampler's camera
// samples are a pure function of (task window, candidate index): candidate i of a task maps to
// image (u, v) = (ri(i, 3), ri(i, 2)) stretched over the window's bounding square and is skipped
// when it falls outside the window; lens and time use the index AFTER the increment (:72-87).
// So every candidate is evaluated independently, one thread each.
//
// What a Halton frame lacks is a fixed number of samples per pixel.  The wavefront buffers are
// therefore padded: `cap` slots per list pixel (cap = the largest per-pixel count), a pixel's
// samples in candidate-index order (one task owns a pixel, so that is the reference's generation
// order restricted to the pixel), unused slots marked by a NaN image coordinate, which the trace,
// and film kernels skip.  Three passes:
//   k_halton_bin<0>  count the accepted candidates per home pixel              (atomics: counts only)
//   k_halton_bin<1>  scatter the candidate indices into the pixel's slots      (arbitrary order)
//   k_halton_samples per pixel: sort its indices, evaluate the camera samples  (deterministic order)
// Light-sample floats are oracle-defined for this sampler: pair q of a camera sample
// is (ri(i + 1, prime[5 + 2q]), ri(i + 1, prime[6 + 2q])) — as written the reference's 1D / 2D
// sample arrays panic (halton.rs:96-107 hands latin_hypercube the slice before the offset).
#pragma once
#include "scene.cuh"

struct DHaltonTask {
  int x0, x1, y0, y1;            // the task's sampler sub-window (sampler/base.rs:29-48)
  float delta;                   // lerp_delta = dy.max(dx) (halton.rs:62-66)
  uint32_t pad;
  unsigned long long first;      // global ordinal of this task's candidate 0
  unsigned long long wanted;     // max(dx, dy)^2 * samples_per_pixel (halton.rs:20-27)
};

#define PB_HALTON_MAX_LIGHT_PAIRS 16
static __device__ const unsigned int pb_halton_primes[40] = {
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

struct HaltonArgs {
  const DHaltonTask* __restrict__ tasks;
  uint32_t n_tasks;
  unsigned long long n_candidates;
  const int32_t* __restrict__ pix_index;  // sampler-extent raster -> list position or -1
  int sx0, sy0, sw;
  uint32_t* counts;   // per list pixel: accepted candidates (pass 0)
  uint32_t* fill;     // per list pixel: slots handed out (pass 1)
  uint32_t* idx;      // [list pixel][cap]: candidate index within its task
  uint32_t cap;
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
    if (s < a.cap) a.idx[(size_t)li * a.cap + s] = (uint32_t)i;
  }
}

// max and sum of the per-pixel counts -> out[0] = max, out[1..2] = sum (u64)
__global__ void __launch_bounds__(256) k_halton_stats(const uint32_t* __restrict__ counts, uint64_t n,
                                                      uint32_t* out_max, unsigned long long* out_sum) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t c = i < n ? counts[i] : 0u;
  uint32_t m = c;
  uint32_t s = c;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (m) atomicMax(out_max, m);
    if (s) atomicAdd(out_sum, (unsigned long long)s);
  }
}

struct HaltonSampleArgs {
  const DHaltonTask* __restrict__ tasks;
  const DPixel* __restrict__ pixels;
  uint64_t n_pixels;
  const uint32_t* __restrict__ counts;
  uint32_t* idx;
  uint32_t cap;
  float2* __restrict__ img;
  float2* __restrict__ lens;    // may be NULL
  float* __restrict__ time;     // may be NULL
  float2* __restrict__ lightu;  // may be NULL; light_pairs float2 per slot
  uint32_t light_pairs;
  uint32_t* __restrict__ edge;  // may be NULL (primary_hits); marked 1: every pixel takes the full gather
  float sopen, sclose;
};

__global__ void __launch_bounds__(128) k_halton_samples(const HaltonSampleArgs a) {
  const uint64_t li = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= a.n_pixels) return;
  const uint32_t cnt = min(a.counts[li], a.cap);
  uint32_t* q = a.idx + li * a.cap;
  for (uint32_t s = 1; s < cnt; ++s) {  // insertion sort: candidate-index order
    const uint32_t v = q[s];
    uint32_t k = s;
    while (k > 0 && q[k - 1] > v) {
      q[k] = q[k - 1];
      --k;
    }
    q[k] = v;
  }
  const DHaltonTask t = a.tasks[a.pixels[li].task & PB_PIXEL_TASK_MASK];
  const uint64_t base = li * a.cap;
  for (uint32_t s = 0; s < a.cap; ++s) {
    float2 im = make_float2(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000));
    float2 ln = make_float2(0.f, 0.f);
    float tm = 0.f;
    unsigned long long cur = 0;
    if (s < cnt) {
      const unsigned long long i = q[s];
      halton_image(t, i, &im.x, &im.y);
      cur = i + 1ull;  // halton.rs:72: the increment precedes the lens / time dimensions
      if (a.lens) ln = make_float2((float)radical_inverse_(cur, 5u), (float)radical_inverse_(cur, 7u));
      if (a.time) tm = lerpf_(a.sopen, a.sclose, (float)radical_inverse_(cur, 11u));
    }
    a.img[base + s] = im;
    if (a.lens) a.lens[base + s] = ln;
    if (a.time) a.time[base + s] = tm;
    if (a.lightu)
      for (uint32_t p = 0; p < a.light_pairs; ++p)
        a.lightu[(base + s) * a.light_pairs + p] =
            s < cnt ? make_float2((float)radical_inverse_(cur, pb_halton_primes[5 + 2 * p]),
                                  (float)radical_inverse_(cur, pb_halton_primes[6 + 2 * p]))
                    : make_float2(0.f, 0.f);
  }
  if (a.edge) a.edge[li] = 1u;
}
