// Camera-sample generation kernels (sm_100a).
//
// Replaces StratifiedSampler::get_more_samples (src/sampler/stratified.rs:60-127) with
// stratified_sample_2d/1d (src/montecarlo.rs:107-129) and RNG::shuffle (src/rng.rs:23-33), and
// LDSampler::get_more_samples (src/sampler/lds.rs:50-70) with ld_pixel_sample & friends
// (src/sampler/utils.rs:6-163).  The per-task StdRng (src/rng.rs:11-13) is a ChaCha12 word stream;
// since every pixel consumes a constant number of words W (DSampler.words_per_pixel), pixel k of a
// task starts at word k*W, which makes the sequential CPU sequence addressable per pixel.
#pragma once
#include "scene.cuh"

// Fast path: stratified sampler, pinhole camera, outputs only (image_x, image_y).
// One thread per camera sample; it derives the one or two ChaCha blocks holding its two words.
__global__ void __launch_bounds__(256)
k_raygen_image(const DSampler smp, const DPixel* __restrict__ pixels, uint64_t n_samples,
               float2* __restrict__ img) {
  const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_samples) return;
  const uint32_t spp = (uint32_t)smp.spp;
  const uint64_t p = s / spp;
  const uint32_t i = (uint32_t)(s - p * spp);
  const DPixel px = pixels[p];
  float jx = 0.5f, jy = 0.5f;
  if (smp.jitter) {
    const uint32_t* key = smp.task_keys + 8u * px.task;
    const uint64_t w = (uint64_t)px.k * smp.words_per_pixel + 2ull * i;
    uint32_t buf[16];
    chacha12_block(key, w >> 4, buf);
    const uint32_t wi = (uint32_t)(w & 15);
    uint32_t a = 0, b = 0;
#pragma unroll
    for (int q = 0; q < 16; ++q) {  // static indexing keeps the block in registers
      if (q == (int)wi) a = buf[q];
      if (q == (int)wi + 1) b = buf[q];
    }
    if (wi == 15) {
      chacha12_block(key, (w >> 4) + 1, buf);
      b = buf[0];
    }
    jx = u32_to_unit_float(a);
    jy = u32_to_unit_float(b);
  }
  const uint32_t sx = i % (uint32_t)smp.xs, sy = i / (uint32_t)smp.xs;
  const float dx = 1.0f / (float)smp.xs, dy = 1.0f / (float)smp.ys;
  float vx = ((float)sx + jx) * dx;  // montecarlo.rs:121-126
  float vy = ((float)sy + jy) * dy;
  vx += (float)px_x(px);  // stratified.rs:79-82
  vy += (float)px_y(px);
  img[s] = make_float2(vx, vy);
}

// sampler/utils.rs:6-20 (as written: the last bit-reversal step shifts by 2)
PB_DEV float van_der_corput_(uint32_t n, uint32_t scramble) {
  n = (n << 16) | (n >> 16);
  n = ((n & 0x00ff00ffu) << 8) | ((n & 0xff00ff00u) >> 8);
  n = ((n & 0x0f0f0f0fu) << 4) | ((n & 0xf0f0f0f0u) >> 4);
  n = ((n & 0x33333333u) << 2) | ((n & 0xCCCCCCCCu) >> 2);
  n = ((n & 0x55555555u) << 2) | ((n & 0xAAAAAAAAu) >> 2);
  n ^= scramble;
  return (float)((double)((n >> 8) & 0xffffffu) / 16777216.0);
}
// sampler/utils.rs:22-35
PB_DEV float sobol2_(uint32_t n, uint32_t scramble) {
  uint32_t s = scramble, v = 1u << 31;
  while (n != 0) {
    if ((n & 1u) == 0) s ^= v;
    v ^= v >> 1;
    n >>= 1;
  }
  return (float)((double)((s >> 8) & 0xFFFFFFu) / 16777216.0);
}

// rng.rs:23-33 over `count` groups of DIMS floats stored as float / float2 in global memory
PB_DEV void shuffle2(WordStream& ws, float2* v, uint32_t count) {
  for (uint32_t i = 0; i < count; ++i) {
    const uint32_t other = i + (uint32_t)(ws.random_uint() % (uint64_t)(count - i));
    const float2 t = v[i];
    v[i] = v[other];
    v[other] = t;
  }
}
PB_DEV void shuffle1(WordStream& ws, float* v, uint32_t count) {
  for (uint32_t i = 0; i < count; ++i) {
    const uint32_t other = i + (uint32_t)(ws.random_uint() % (uint64_t)(count - i));
    const float t = v[i];
    v[i] = v[other];
    v[other] = t;
  }
}

// General path: one thread per pixel runs the pixel's whole sample block, including the lens /
// time shuffles and the LD sampler; results are built in place in the thread's own slice of the
// output arrays.  `time` receives lerp(shutter_open, shutter_close, t) (stratified.rs:92-93).
__global__ void __launch_bounds__(128)
k_raygen_full(const DSampler smp, const DPixel* __restrict__ pixels, uint64_t n_pixels,
              float2* __restrict__ img, float2* __restrict__ lens, float* __restrict__ time) {
  const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pixels) return;
  const DPixel px = pixels[p];
  const uint32_t n = (uint32_t)smp.spp;
  float2* im = img + p * n;
  float2* ln = lens + p * n;
  float* tm = time + p * n;
  WordStream ws;
  ws.init(smp.task_keys + 8u * px.task, (uint64_t)px.k * smp.words_per_pixel);
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
  for (uint32_t i = 0; i < n; ++i) tm[i] = lerpf_(smp.sopen, smp.sclose, tm[i]);
}
