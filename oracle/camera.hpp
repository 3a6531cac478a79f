// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// Filter, Film, perspective Camera and the Stratified / LowDiscrepancy samplers of pbrt_rust,
// restated line-faithfully (citations relative to the reference root).
#pragma once
#include <algorithm>

#include "geom.hpp"
#include "rng.hpp"

namespace orc {

// utils/mod.rs:207-218
inline float sinc_1d(float x, float tau) {
  float v = std::fabs(x);
  if (v < 1e-5f) return 1.0f;
  if (v >= 1.0f) return 0.0f;
  v *= PI_F;
  float vtau = v * tau;
  float s = std::sin(vtau) / vtau;
  return s * std::sin(v) / v;
}

// filter.rs
struct Filter {
  enum Ty { Mean = 0, Triangle = 1, Gaussian = 2, Mitchell = 3, Lanczos = 4 } ty = Mean;
  float xw = 0.5f, yw = 0.5f, inv_xw = 2.f, inv_yw = 2.f;
  float p0 = 0.f, p1 = 0.f;  // Gaussian: alpha ; Mitchell: b, c ; Lanczos: tau
  float exp_x = 0.f, exp_y = 0.f;
  Filter() = default;
  Filter(int type, float xw_, float yw_, float a, float b)
      : ty((Ty)type), xw(xw_), yw(yw_), inv_xw(1.0f / xw_), inv_yw(1.0f / yw_), p0(a), p1(b) {
    if (ty == Gaussian) {  // filter.rs:60-69
      exp_x = std::exp(-a * xw_ * xw_);
      exp_y = std::exp(-a * yw_ * yw_);
    }
  }
  // filter.rs:87-124
  float evaluate(float x, float y) const {
    switch (ty) {
      case Mean:
        return 1.0f;
      case Triangle: {
        float dx = (xw - std::fabs(x)) * inv_xw;
        float dy = (yw - std::fabs(y)) * inv_yw;
        return rmax(dx, 0.0f) * rmax(dy, 0.0f);
      }
      case Gaussian: {
        auto g = [&](float v, float ex) { return rmax(std::exp(-p0 * v * v) - ex, 0.0f); };
        return g(x, exp_x) * g(y, exp_y);
      }
      case Mitchell: {
        float b = p0, c = p1;
        auto mitchell = [&](float v) {
          float t = std::fabs(v * 2.0f);
          float r;
          if (t >= 2.0f)
            r = 0.0f;
          else if (t > 1.0f)
            r = (-b - 6.0f * c) * t * t * t + (6.0f * b + 30.0f * c) * t * t +
                (-12.0f * b - 48.0f * c) * t + (8.0f * b + 24.0f * c);
          else
            r = (12.0f - 9.0f * b - 6.0f * c) * t * t * t +
                (-18.0f + 12.0f * b + 6.0f * c) * t * t + (6.0f - 2.0f * b);
          return (1.0f / 6.0f) * r;
        };
        return mitchell(x * inv_xw) * mitchell(y * inv_yw);
      }
      case Lanczos:
        return sinc_1d(x * inv_xw, p0) * sinc_1d(y * inv_yw, p0);
    }
    return 0.f;
  }
};

static constexpr int FILTER_TABLE_DIM = 16;  // film.rs:12

struct CameraSample {  // camera/mod.rs:19-26
  float image_x = 0.f, image_y = 0.f, lens_u = 0.f, lens_v = 0.f, time = 0.f;
};

struct Pixel {  // film.rs:35-41
  float xyz[3] = {0.f, 0.f, 0.f};
  float weight_sum = 0.f;
};

// spectrum.rs:37-41
inline void rgb_to_xyz(const float rgb[3], float xyz[3]) {
  xyz[0] = 0.412453f * rgb[0] + 0.357580f * rgb[1] + 0.180423f * rgb[2];
  xyz[1] = 0.212671f * rgb[0] + 0.715160f * rgb[1] + 0.072169f * rgb[2];
  xyz[2] = 0.019334f * rgb[0] + 0.119193f * rgb[1] + 0.950227f * rgb[2];
}
// spectrum.rs:31-35
inline void xyz_to_rgb(const float xyz[3], float rgb[3]) {
  rgb[0] = 3.240479f * xyz[0] - 1.37150f * xyz[1] - 0.498535f * xyz[2];
  rgb[1] = -0.969256f * xyz[0] + 1.875991f * xyz[1] + 0.041556f * xyz[2];
  rgb[2] = 0.055648f * xyz[0] - 0.204043f * xyz[1] + 1.057311f * xyz[2];
}

// camera/film.rs:69-309
struct Film {
  int x_res = 0, y_res = 0;
  Filter filter;
  float crop[4] = {0.f, 1.f, 0.f, 1.f};
  int x_start = 0, y_start = 0, x_count = 0, y_count = 0;
  std::vector<Pixel> pixels;
  float table[FILTER_TABLE_DIM * FILTER_TABLE_DIM];

  Film() = default;
  // film.rs:77-122
  Film(int xres, int yres, const Filter& f, const float c[4]) : x_res(xres), y_res(yres), filter(f) {
    for (int i = 0; i < 4; ++i) crop[i] = c[i];
    x_start = f2i(std::ceil((float)xres * c[0]));
    x_count = std::max(f2i(std::ceil((float)xres * c[1])) - x_start, 1);
    y_start = f2i(std::ceil((float)yres * c[2]));
    y_count = std::max(f2i(std::ceil((float)yres * c[3])) - y_start, 1);
    for (int y = 0; y < FILTER_TABLE_DIM; ++y) {
      float fy = ((float)y + 0.5f) * f.yw / (float)FILTER_TABLE_DIM;
      for (int x = 0; x < FILTER_TABLE_DIM; ++x) {
        float fx = ((float)x + 0.5f) * f.xw / (float)FILTER_TABLE_DIM;
        table[y * FILTER_TABLE_DIM + x] = f.evaluate(fx, fy);
      }
    }
    pixels.assign((size_t)x_count * (size_t)y_count, Pixel());
  }
  // film.rs:124-147
  Film sub_film(size_t num, size_t count) const {
    float aspect = (float)x_res / (float)y_res;
    float w[4];
    get_crop_window(num, count, aspect, w);
    float dx = crop[1] - crop[0];
    float dy = crop[3] - crop[2];
    float cw[4] = {crop[0] + w[0] * dx, crop[0] + w[1] * dx, crop[2] + w[2] * dy,
                   crop[2] + w[3] * dy};
    return Film(x_res, y_res, filter, cw);
  }
  // film.rs:271-289
  void sample_extent(int out[4]) const {
    out[0] = f2i(std::floor((float)x_start + 0.5f - filter.xw));
    out[1] = f2i(std::floor((float)x_start + 0.5f + (float)x_count + filter.xw));
    out[2] = f2i(std::floor((float)y_start + 0.5f - filter.yw));
    out[3] = f2i(std::floor((float)y_start + 0.5f + (float)y_count + filter.yw));
  }
  // film.rs:294-303
  void pixel_extent(int out[4]) const {
    out[0] = x_start;
    out[1] = x_start + x_count;
    out[2] = y_start;
    out[3] = y_start + y_count;
  }
  // film.rs:192-249.  `rgb` = the RGB spectrum L.
  void add_sample(const CameraSample& s, const float rgb[3]) {
    float dimage_x = s.image_x - 0.5f;
    float dimage_y = s.image_y - 0.5f;
    int x0 = std::max(x_start, f2i(std::ceil(dimage_x - filter.xw)));
    int x1 = std::min(x_start + x_count - 1, f2i(std::floor(dimage_x + filter.xw)));
    int y0 = std::max(y_start, f2i(std::ceil(dimage_y - filter.yw)));
    int y1 = std::min(y_start + y_count - 1, f2i(std::floor(dimage_y + filter.yw)));
    if ((x1 - x0) < 0 || (y1 - y0) < 0) return;
    float xyz[3];
    rgb_to_xyz(rgb, xyz);
    for (int y = y0; y <= y1; ++y) {
      float fy = ((float)y - dimage_y) * filter.inv_yw * (float)FILTER_TABLE_DIM;
      size_t iy = std::min<size_t>((size_t)f2usize(std::floor(std::fabs(fy))), FILTER_TABLE_DIM - 1);
      for (int x = x0; x <= x1; ++x) {
        float fx = ((float)x - dimage_x) * filter.inv_xw * (float)FILTER_TABLE_DIM;
        size_t ix =
            std::min<size_t>((size_t)f2usize(std::floor(std::fabs(fx))), FILTER_TABLE_DIM - 1);
        float wt = table[iy * FILTER_TABLE_DIM + ix];
        Pixel& px = pixels[(size_t)(y - y_start) * (size_t)x_count + (size_t)(x - x_start)];
        px.xyz[0] += wt * xyz[0];
        px.xyz[1] += wt * xyz[1];
        px.xyz[2] += wt * xyz[2];
        px.weight_sum += wt;
      }
    }
  }
  // film.rs:149-186  (assign_pixels overwrites, D13)
  void add_sub_film(const Film& f) {
    int l[4], g[4];
    f.pixel_extent(l);
    pixel_extent(g);
    if (l[0] < g[0] || l[2] < g[2] || l[1] > g[1] || l[3] > g[3])
      throw std::runtime_error("assign_pixels: sub film outside film");
    int lstride = l[1] - l[0], gstride = g[1] - g[0];
    for (int x = l[0]; x < l[1]; ++x)
      for (int y = l[2]; y < l[3]; ++y)
        pixels[(size_t)((y - g[2]) * gstride + (x - g[0]))] =
            f.pixels[(size_t)((y - l[2]) * lstride + (x - l[0]))];
  }
  // D6: the evident intent of film.rs:311-355 (the as-written code panics):
  // rgb = max(0, xyz_to_rgb(xyz) * (1/weight_sum)) when weight_sum != 0.
  void to_rgb(float* out) const {
    for (size_t i = 0; i < pixels.size(); ++i) {
      float rgb[3];
      xyz_to_rgb(pixels[i].xyz, rgb);
      float w = pixels[i].weight_sum;
      if (w != 0.0f) {
        float inv = 1.0f / w;
        rgb[0] = rmax(rgb[0] * inv, 0.0f);
        rgb[1] = rmax(rgb[1] * inv, 0.0f);
        rgb[2] = rmax(rgb[2] * inv, 0.0f);
      }
      out[3 * i] = rgb[0];
      out[3 * i + 1] = rgb[1];
      out[3 * i + 2] = rgb[2];
    }
  }
};

// camera/mod.rs:105-135 + camera/projective.rs:48-72, static camera-to-world only.
struct PerspectiveCamera {
  Transform cam_to_world;
  Transform raster_to_camera;
  V3 dx_camera, dy_camera;
  float shutter_open = 0.f, shutter_close = 0.f;
  float lens_radius = 0.f, focal_distance = 0.f;

  PerspectiveCamera() = default;
  PerspectiveCamera(const Transform& c2w, const float sw[4], float sopen, float sclose, float lensr,
                    float focald, float fov, int x_res, int y_res)
      : cam_to_world(c2w),
        shutter_open(sopen),
        shutter_close(sclose),
        lens_radius(lensr),
        focal_distance(focald) {
    float znear = 1e-2f, zfar = 1000.0f;
    M44 pm = M44::rows(1, 0, 0, 0, 0, 1, 0, 0, 0, 0, zfar / (zfar - znear),
                       -(zfar * znear) / (zfar - znear), 0, 0, 1, 0);
    Transform p = Transform::from_matrix(pm);
    float inv_tan_ang = 1.0f / std::tan(as_radians(fov) / 2.0f);
    Transform persp = Transform::scale(inv_tan_ang, inv_tan_ang, 1.0f) * p;
    // projective.rs:55-61
    Transform screen_to_raster =
        Transform::scale((float)x_res, (float)y_res, 1.0f) *
        Transform::scale(1.0f / (sw[1] - sw[0]), 1.0f / (sw[2] - sw[3]), 1.0f) *
        Transform::translate(V3(-sw[0], -sw[3], 0.0f));
    Transform raster_to_screen = screen_to_raster.inverse();
    raster_to_camera = persp.inverse() * raster_to_screen;
    // camera/mod.rs:129-133 (D16: Vector transforms, not Point differences)
    dx_camera = raster_to_camera.vec(V3(1, 0, 0)) - raster_to_camera.vec(V3(0, 0, 0));
    dy_camera = raster_to_camera.vec(V3(0, 1, 0)) - raster_to_camera.vec(V3(0, 0, 0));
  }
  // projective.rs:79-97 (concentric_sample_disk is the identity, D14)
  void handle_dof(const CameraSample& s, Ray& ray) const {
    if (lens_radius <= 0.0f) return;
    float u = s.lens_u, v = s.lens_v;
    u *= lens_radius;
    v *= lens_radius;
    float ft = focal_distance / ray.d.z;
    V3 p_focus = ray.at(ft);
    ray.o = V3(u, v, 0.0f);
    ray.d = normalize(p_focus - ray.o);
  }
  // camera/mod.rs:212-271 (Perspective arm) ; weight is always 1.
  RayDifferential generate_ray_differential(const CameraSample& s) const {
    RayDifferential rd;
    rd.has_differentials = true;
    V3 p_camera = raster_to_camera.pt(V3(s.image_x, s.image_y, 0.0f));
    rd.ray = Ray(V3(), normalize(p_camera), 0.0f);  // mod.rs:168-180
    rd.rx_origin = rd.ray.o;
    rd.ry_origin = rd.ray.o;
    rd.rx_dir = normalize(p_camera + dx_camera);
    rd.ry_dir = normalize(p_camera + dy_camera);
    handle_dof(s, rd.ray);
    rd.ray.time = lerpf(shutter_open, shutter_close, s.time);
    // animated.rs:275-284 with a static transform: only ray.o / ray.d move (D15).
    rd.ray.o = cam_to_world.pt(rd.ray.o);
    rd.ray.d = cam_to_world.vec(rd.ray.d);
    return rd;
  }
};

// ---------------------------------------------------------------------------------------------
// sampler/base.rs:29-48
inline void compute_sub_window(const int ext[4], size_t num, size_t count, int out[4]) {
  size_t dx = (size_t)(ext[1] - ext[0]);
  size_t dy = (size_t)(ext[3] - ext[2]);
  float aspect = (float)dx / (float)dy;
  float t[4];
  get_crop_window(num, count, aspect, t);
  float psx = (float)ext[0], psy = (float)ext[2], pex = (float)ext[1], pey = (float)ext[3];
  out[0] = f2i(lerpf(psx, pex, t[0]));
  out[1] = f2i(lerpf(psx, pex, t[1]));
  out[2] = f2i(lerpf(psy, pey, t[2]));
  out[3] = f2i(lerpf(psy, pey, t[3]));
}

// sampler/utils.rs:6-20 (D18: last bit-reversal step shifts by 2)
inline float van_der_corput(uint32_t n, uint32_t scramble) {
  n = (n << 16) | (n >> 16);
  n = ((n & 0x00ff00ffu) << 8) | ((n & 0xff00ff00u) >> 8);
  n = ((n & 0x0f0f0f0fu) << 4) | ((n & 0xf0f0f0f0u) >> 4);
  n = ((n & 0x33333333u) << 2) | ((n & 0xCCCCCCCCu) >> 2);
  n = ((n & 0x55555555u) << 2) | ((n & 0xAAAAAAAAu) >> 2);
  n ^= scramble;
  return (float)((double)((n >> 8) & 0xffffffu) / (double)(1 << 24));
}
// sampler/utils.rs:22-35
inline float sobol2(uint32_t n, uint32_t scramble) {
  uint32_t s = scramble;
  uint32_t v = 1u << 31;
  while (n != 0) {
    if ((n & 0x1u) == 0) s ^= v;
    v ^= v >> 1;
    n >>= 1;
  }
  return (float)((double)((s >> 8) & 0xFFFFFFu) / (double)(1 << 24));
}
// sampler/utils.rs:49-63
inline void ld_shuffle_scrambled_1d(size_t num_samples, size_t num_pixel_samples, float* samples,
                                    size_t len, RNG& rng) {
  uint32_t scramble = (uint32_t)rng.random_uint();
  for (size_t i = 0; i < num_samples * num_pixel_samples; ++i)
    samples[i] = van_der_corput((uint32_t)i, scramble);
  for (size_t off = 0; off < len; off += num_samples)
    rng.shuffle(samples + off, std::min(num_samples, len - off), 1);
  rng.shuffle(samples, len, num_samples);
}
// sampler/utils.rs:65-83
inline void ld_shuffle_scrambled_2d(size_t num_samples, size_t num_pixel_samples, float* samples,
                                    size_t len, RNG& rng) {
  uint32_t sc0 = (uint32_t)rng.random_uint();
  uint32_t sc1 = (uint32_t)rng.random_uint();
  for (size_t i = 0; i < num_samples * num_pixel_samples; ++i) {
    samples[2 * i] = van_der_corput((uint32_t)i, sc0);
    samples[2 * i + 1] = sobol2((uint32_t)i, sc1);
  }
  for (size_t off = 0; off < len; off += 2 * num_samples)
    rng.shuffle(samples + off, std::min(2 * num_samples, len - off), 2);
  rng.shuffle(samples, len, 2 * num_samples);
}

// A sampler restricted to one (sub-)window.  kind 0 = Stratified (sampler/stratified.rs),
// kind 1 = LowDiscrepancy (sampler/lds.rs).
struct SamplerDesc {
  int kind = 0;
  int ext[4] = {0, 0, 0, 0};  // x_start, x_end, y_start, y_end
  int xs = 1, ys = 1;         // stratified strata ; LD: xs = spp (rounded up to pow2), ys = 1
  bool jitter = true;
  float sopen = 0.f, sclose = 0.f;
  // Oracle-defined extension: `light_samples` pairs of floats per camera sample,
  // drawn from the same stream AFTER the pixel's camera-sample block.
  int light_samples = 0;

  size_t spp() const {
    if (kind == 0) return (size_t)xs * (size_t)ys;
    if (kind == 2) return (size_t)xs;  // HaltonSampler: samples_per_pixel as given (halton.rs:18-29)
    size_t p = 1;
    while (p < (size_t)xs) p <<= 1;  // lds.rs:18 next_power_of_two
    return p;
  }
  // RNG words consumed per pixel (SURVEY §8a A1/A1').
  size_t words_per_pixel() const {
    size_t n = spp();
    size_t cam = (kind == 0) ? (jitter ? 9 * n : 4 * n) : 2 * (5 + 6 * n);
    return cam + 2 * n * (size_t)light_samples;
  }
};

// ---------------------------------------------------------------------------------------------
// HaltonSampler (sampler/halton.rs), sampler kind 2.
// montecarlo.rs:7-20 (f64 arithmetic, as written)
inline double radical_inverse(uint64_t n, uint64_t b) {
  double v = 0.0;
  uint64_t num = n;
  const double inv_base = 1.0 / (double)b;
  double aib = 1.0;
  while (num > 0) {
    double d = (double)(num % b);
    num /= b;
    aib *= inv_base;
    v += d * aib;
  }
  return v;
}
static const uint32_t HALTON_PRIMES[40] = {2,  3,  5,  7,  11, 13, 17, 19, 23, 29,  31,  37,  41,  43,
                                           47, 53, 59, 61, 67, 71, 73, 79, 83, 89,  97,  101, 103, 107,
                                           109, 113, 127, 131, 137, 139, 149, 151, 157, 163, 167, 173};
#define ORC_HALTON_MAX_LIGHT_PAIRS 16
// One (sub-)sampler: HaltonSampler::new over a window (halton.rs:18-29)
struct HaltonWindow {
  int ext[4];
  float delta;      // lerp_delta = dy.max(dx) (halton.rs:62-66)
  uint64_t wanted;  // max(dx, dy)^2 * samples_per_pixel
};
inline HaltonWindow halton_window(const int ext[4], size_t spp) {
  HaltonWindow w;
  for (int i = 0; i < 4; ++i) w.ext[i] = ext[i];
  const int dx = ext[1] - ext[0], dy = ext[3] - ext[2];
  const int m = dx > dy ? dx : dy;
  w.wanted = (uint64_t)((int64_t)m * (int64_t)m) * (uint64_t)spp;
  w.delta = rmax((float)dy, (float)dx);
  return w;
}
// get_more_samples for candidate index i = current_sample (halton.rs:49-108).  Returns false when
// the candidate falls outside the window and is skipped.  The lens / time dimensions use the index
// AFTER the increment (`self.current_sample += 1` precedes them, :72), as written.
// Light-sample floats: ORACLE-DEFINED (as written the 1D/2D sample arrays panic —
// `split_at_mut(off)` hands latin_hypercube the slice before the offset, :96-107): pair q of the
// camera sample = radical inverses of the incremented index in bases prime[5 + 2q], prime[6 + 2q].
inline bool halton_candidate(const SamplerDesc& sd, const HaltonWindow& w, uint64_t i, CameraSample* cs,
                             float* light_u) {
  const float u = (float)radical_inverse(i, 3);
  const float v = (float)radical_inverse(i, 2);
  const float xs = (float)w.ext[0], ys = (float)w.ext[2];
  const float image_x = lerpf(xs, xs + w.delta, u);
  const float image_y = lerpf(ys, ys + w.delta, v);
  const uint64_t cur = i + 1;
  if (image_x >= (float)w.ext[1] || image_y >= (float)w.ext[3]) return false;
  cs->image_x = image_x;
  cs->image_y = image_y;
  cs->lens_u = (float)radical_inverse(cur, 5);
  cs->lens_v = (float)radical_inverse(cur, 7);
  cs->time = lerpf(sd.sopen, sd.sclose, (float)radical_inverse(cur, 11));
  for (int q = 0; q < sd.light_samples; ++q) {
    light_u[2 * q] = (float)radical_inverse(cur, HALTON_PRIMES[5 + 2 * q]);
    light_u[2 * q + 1] = (float)radical_inverse(cur, HALTON_PRIMES[6 + 2 * q]);
  }
  return true;
}
// Home pixel of an accepted sample inside its window (binning only: add_sample decides coverage).
inline void halton_home_pixel(const HaltonWindow& w, const CameraSample& cs, int* px, int* py) {
  int x = f2i(std::floor(cs.image_x)), y = f2i(std::floor(cs.image_y));
  *px = x < w.ext[0] ? w.ext[0] : (x > w.ext[1] - 1 ? w.ext[1] - 1 : x);
  *py = y < w.ext[2] ? w.ext[2] : (y > w.ext[3] - 1 ? w.ext[3] - 1 : y);
}

// Generates the `spp` camera samples (and light-sample floats) of ONE pixel, advancing `rng`
// exactly as the reference's get_more_samples does (stratified.rs:60-127 / lds.rs:50-70 +
// sampler/utils.rs:85-163 with Sample::empty(), i.e. no integrator-requested samples).
inline void pixel_samples(const SamplerDesc& sd, int x_pos, int y_pos, RNG& rng,
                          std::vector<CameraSample>& out, std::vector<float>& light_u) {
  size_t n = sd.spp();
  out.resize(n);
  std::vector<float> buf(5 * n);
  float* image = buf.data();
  float* lens = image + 2 * n;
  float* time = lens + 2 * n;
  if (sd.kind == 0) {
    stratified_sample_2d(image, (size_t)sd.xs, (size_t)sd.ys, rng, sd.jitter);
    stratified_sample_2d(lens, (size_t)sd.xs, (size_t)sd.ys, rng, sd.jitter);
    stratified_sample_1d(time, n, rng, sd.jitter);
    for (size_t i = 0; i < n; ++i) {
      image[2 * i] += (float)x_pos;
      image[2 * i + 1] += (float)y_pos;
    }
    rng.shuffle(lens, 2 * n, 2);
    rng.shuffle(time, n, 1);
    for (size_t i = 0; i < n; ++i) {
      out[i].image_x = image[2 * i];
      out[i].image_y = image[2 * i + 1];
      out[i].lens_u = lens[2 * i];
      out[i].lens_v = lens[2 * i + 1];
      out[i].time = lerpf(sd.sopen, sd.sclose, time[i]);
    }
  } else {
    ld_shuffle_scrambled_2d(1, n, image, 2 * n, rng);
    ld_shuffle_scrambled_2d(1, n, lens, 2 * n, rng);
    ld_shuffle_scrambled_1d(1, n, time, n, rng);
    for (size_t i = 0; i < n; ++i) {
      out[i].image_x = (float)x_pos + image[2 * i];
      out[i].image_y = (float)y_pos + image[2 * i + 1];
      out[i].lens_u = lens[2 * i];
      out[i].lens_v = lens[2 * i + 1];
      out[i].time = lerpf(sd.sopen, sd.sclose, time[i]);
    }
  }
  light_u.resize(2 * n * (size_t)sd.light_samples);
  for (size_t i = 0; i < light_u.size(); ++i) light_u[i] = rng.random_float();
}

// sampler_renderer.rs:37-47 (D12): ceil(log2(max(32*ncpu, npix/256)))
inline uint32_t num_tasks_for(uint32_t num_cpus, uint32_t num_pixels) {
  uint32_t x = std::max(32u * num_cpus, num_pixels / 256u);
  uint32_t lz = (uint32_t)__builtin_clz(x);
  return 31u - lz + ((x & (x - 1)) == 0 ? 0u : 1u);
}

}  // namespace orc
