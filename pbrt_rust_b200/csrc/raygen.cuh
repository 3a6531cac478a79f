// Camera-sample generation kernels (sm_100a).
//
// Replaces StratifiedSampler::get_more_samples (src/sampler/stratified.rs:60-127) with
// stratified_sample_2d/1d (src/montecarlo.rs:107-129) and RNG::shuffle (src/rng.rs:23-33), and
// LDSampler::get_more_samples (src/sampler/lds.rs:50-70) with ld_pixel_sample & friends
// (src/sampler/utils.rs:6-163).  The per-task StdRng (src/rng.rs:11-13) is a ChaCha12 word stream;
// since every pixel consumes a constant number of words W (DSampler.words_per_pixel), pixel k of a
// task starts at word k*W, which makes the sequential CPU sequence addressable per pixel.
//
// Outputs (list order, sample s = pixel_pos * spp + i):
//   img[s]            (image_x, image_y)
//   lens[s], time[s]  only from the general kernel (depth of field / out_samples requested)
//   lightu[s*P + j]   the j-th light-sample float pair of camera sample s (SURVEY D11 extension)
//   edge[pixel_pos]   != 0 iff some sample of the pixel has an add_sample extent (film.rs:198-210)
//                     other than exactly its own pixel — lets the film gather skip neighbours.
#pragma once
#include "scene.cuh"

struct RaygenArgs {
  const DPixel* __restrict__ pixels;
  uint64_t n_pixels;
  float2* __restrict__ img;
  float2* __restrict__ lens;
  float* __restrict__ time;
  float2* __restrict__ lightu;  // NULL when the scene has no area lights
  uint32_t* __restrict__ edge;  // per list pixel, zeroed by the host before the frame
  uint32_t light_pairs;
  // film geometry for the edge flags
  int fx_start, fy_start, fx_count, fy_count;
  float xw, yw;
};

// film.rs:198-210 evaluated for one sample: does it touch any pixel other than (px, py)?
PB_DEV bool sample_leaves_own_pixel(const RaygenArgs& a, float ix, float iy, int px, int py) {
  const float dimage_x = ix - 0.5f, dimage_y = iy - 0.5f;
  const int x0 = f2i_sat(ceilf(dimage_x - a.xw)), x1 = f2i_sat(floorf(dimage_x + a.xw));
  const int y0 = f2i_sat(ceilf(dimage_y - a.yw)), y1 = f2i_sat(floorf(dimage_y + a.yw));
  return !(x0 == px && x1 == px && y0 == py && y1 == py);
}

// 16 consecutive stream words starting at absolute word `w0` (one block if aligned, else two).
PB_DEV void fetch16(const uint32_t* __restrict__ key, uint64_t w0, uint32_t out[16]) {
  const uint32_t r = (uint32_t)(w0 & 15);
  if (r == 0) {
    chacha12_block(key, w0 >> 4, out);
    return;
  }
  uint32_t b0[16], b1[16];
  chacha12_block(key, w0 >> 4, b0);
  chacha12_block(key, (w0 >> 4) + 1, b1);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    // out[j] = word r + j of the 32-word window (b0 | b1); select with static indices
    uint32_t v = 0;
#pragma unroll
    for (int q = 0; q < 32; ++q)
      if (q == (int)r + j) v = q < 16 ? b0[q & 15] : b1[q & 15];
    out[j] = v;
  }
}

// Fast path: stratified sampler, pinhole camera.  One thread per (pixel, group of 8 samples): it
// derives the ChaCha block(s) holding the group's 16 image-jitter words once (instead of once per
// sample) and likewise the group's light-sample words.
__global__ void __launch_bounds__(128)
k_raygen_groups(const DSampler smp, const RaygenArgs a) {
  const uint32_t spp = (uint32_t)smp.spp;
  const uint32_t groups = (spp + 7u) / 8u;
  const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= a.n_pixels * groups) return;
  const uint64_t p = gid / groups;
  const uint32_t g = (uint32_t)(gid - p * groups);
  const DPixel px = a.pixels[p];
  const uint32_t* key = smp.task_keys + 8u * (px.task & PB_PIXEL_TASK_MASK);
  const uint64_t base = (uint64_t)px.k * smp.words_per_pixel;
  const uint32_t i0 = 8u * g, cnt = min(8u, spp - i0);
  const int pxx = px_x(px), pxy = px_y(px);
  const float dx = 1.0f / (float)smp.xs, dy = 1.0f / (float)smp.ys;
  uint32_t w[16];
  if (smp.jitter) fetch16(key, base + 2ull * i0, w);
  bool edge = false;
#pragma unroll
  for (uint32_t j = 0; j < 8; ++j) {
    if (j >= cnt) break;
    const uint32_t i = i0 + j;
    const float jx = smp.jitter ? u32_to_unit_float(w[2 * j]) : 0.5f;
    const float jy = smp.jitter ? u32_to_unit_float(w[2 * j + 1]) : 0.5f;
    const uint32_t sx = i % (uint32_t)smp.xs, sy = i / (uint32_t)smp.xs;
    float vx = ((float)sx + jx) * dx;  // montecarlo.rs:121-126
    float vy = ((float)sy + jy) * dy;
    vx += (float)pxx;  // stratified.rs:79-82
    vy += (float)pxy;
    st_stream(a.img + (p * spp + i), make_float2(vx, vy));
    if (a.edge) edge |= sample_leaves_own_pixel(a, vx, vy, pxx, pxy);
  }
  if (edge && a.edge) atomicOr(&a.edge[p], 1u);
  if (a.lightu) {
    const uint32_t P = a.light_pairs;
    // light floats of sample i start at cam_words + 2*P*i: the group's 8*P pairs are contiguous
    for (uint32_t q = 0; q < P; ++q) {  // 16 words = 8 pairs per fetch
      fetch16(key, base + smp.cam_words + 2ull * P * i0 + 16ull * q, w);
#pragma unroll
      for (uint32_t j = 0; j < 8; ++j) {
        const uint32_t pair = 8u * q + j;  // pair index within the group's 8*P pairs
        if (pair >= cnt * P) break;
        a.lightu[(p * spp + i0) * P + pair] =
            make_float2(u32_to_unit_float(w[2 * j]), u32_to_unit_float(w[2 * j + 1]));
      }
    }
  }
}

// General path: one thread per pixel runs the pixel's whole sample block, including the lens /
// time shuffles and the LD sampler; results are built in place in the thread's own slice of the
// output arrays.  `time` receives lerp(shutter_open, shutter_close, t) (stratified.rs:92-93).
__global__ void __launch_bounds__(128)
k_raygen_full(const DSampler smp, const RaygenArgs a) {
  const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.n_pixels) return;
  const DPixel px = a.pixels[p];
  const uint32_t n = (uint32_t)smp.spp;
  float2* im = a.img + p * n;
  float2* ln = a.lens + p * n;
  float* tm = a.time + p * n;
  WordStream ws;
  ws.init(smp.task_keys + 8u * (px.task & PB_PIXEL_TASK_MASK), (uint64_t)px.k * smp.words_per_pixel);
  const float xpos = (float)px_x(px), ypos = (float)px_y(px);
  if (smp.kind == 0) {
    const uint32_t nx = (uint32_t)smp.xs, ny = (uint32_t)smp.ys;
    const float dx = 1.0f / (float)nx, dy = 1.0f / (float)ny;
    const bool jit = smp.jitter != 0;
    for (uint32_t y = 0; y < ny; ++y)
      for (uint32_t x = 0; x < nx; ++x) {
        const float jx = jit ? ws.random_float() : 0.5f;
        const float jy = jit ? ws.random_float() : 0.5f;
        float vx = ((float)x + jx) * dx, vy = ((float)y + jy) * dy;
        vx += xpos;
        vy += ypos;
        im[y * nx + x] = make_float2(vx, vy);
      }
    for (uint32_t y = 0; y < ny; ++y)
      for (uint32_t x = 0; x < nx; ++x) {
        const float jx = jit ? ws.random_float() : 0.5f;
        const float jy = jit ? ws.random_float() : 0.5f;
        ln[y * nx + x] = make_float2(((float)x + jx) * dx, ((float)y + jy) * dy);
      }
    const float inv_tot = 1.0f / (float)n;
    for (uint32_t i = 0; i < n; ++i) {
      const float delta = jit ? ws.random_float() : 0.5f;
      tm[i] = ((float)i + delta) * inv_tot;
    }
    shuffle2(ws, ln, n);  // stratified.rs:85-86
    shuffle1(ws, tm, n);
  } else {
    // ld_pixel_sample with Sample::empty(): image 2D, lens 2D, time 1D (sampler/utils.rs:123-126)
    {
      const uint32_t sc0 = (uint32_t)ws.random_uint(), sc1 = (uint32_t)ws.random_uint();
      for (uint32_t i = 0; i < n; ++i) im[i] = make_float2(van_der_corput_(i, sc0), sobol2_(i, sc1));
      for (uint32_t i = 0; i < n; ++i) (void)ws.random_uint();  // per-chunk shuffles of 1 group
      shuffle2(ws, im, n);
    }
    {
      const uint32_t sc0 = (uint32_t)ws.random_uint(), sc1 = (uint32_t)ws.random_uint();
      for (uint32_t i = 0; i < n; ++i) ln[i] = make_float2(van_der_corput_(i, sc0), sobol2_(i, sc1));
      for (uint32_t i = 0; i < n; ++i) (void)ws.random_uint();
      shuffle2(ws, ln, n);
    }
    {
      const uint32_t sc = (uint32_t)ws.random_uint();
      for (uint32_t i = 0; i < n; ++i) tm[i] = van_der_corput_(i, sc);
      for (uint32_t i = 0; i < n; ++i) (void)ws.random_uint();
      shuffle1(ws, tm, n);
    }
    for (uint32_t i = 0; i < n; ++i) {  // sampler/utils.rs:131-135: x_pos as f32 + image_samples
      const float2 v = im[i];
      im[i] = make_float2(xpos + v.x, ypos + v.y);
    }
  }
  bool edge = false;
  for (uint32_t i = 0; i < n; ++i) {
    tm[i] = lerpf_(smp.sopen, smp.sclose, tm[i]);
    if (a.edge) edge |= sample_leaves_own_pixel(a, im[i].x, im[i].y, px_x(px), px_y(px));
  }
  if (edge && a.edge) a.edge[p] = 1u;
  if (a.lightu) {  // the stream position is now exactly base + cam_words
    const uint32_t tot = n * a.light_pairs;
    for (uint32_t q = 0; q < tot; ++q) {
      const float u1 = ws.random_float();
      const float u2 = ws.random_float();
      st_stream(a.lightu + (p * tot + q), make_float2(u1, u2));
    }
  }
}
