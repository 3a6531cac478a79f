// TEST INFRASTRUCTURE ONLY — part of the CPU oracle (see oracle/README or geom.hpp header).
// Line-faithful restatement of the reference's image texture path:
//   texture/mipmap.rs   (resample_weights :22-43, resize_to_power_of_two_dims :45-106, ulog2,
//                        texel_at :112-138, MIPMap::new :159-204, triangle :212-226,
//                        pyramid_lookup :228-241, ewa :243-300, lookup :302-341)
//   texture/imagemap.rs (read_image's byte/255 conversion :75-89, TextureCache::get_texture's
//                        (s * scale).powf(gamma) / (s.y() * scale).powf(gamma) :96-173)
//   utils/mod.rs        (sinc_1d :207-217, modulo :219-223, Lerp::lerp_with :20)
// The texel type is always RGB here: a float texture is evaluated per channel with the same
// operations (Mul<f32>, Add, Div<f32>, Sum), so three equal channels reproduce it exactly.
// Pinned by the reference's own tests in imagemap.rs:211-418 (tests/test_oracle_kat.py).
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#include "shading.hpp"  // RGB, f2i, rmax (this header is included from the middle of shading.hpp)

namespace orc {

enum ImageWrap : int { WRAP_REPEAT = 0, WRAP_BLACK = 1, WRAP_CLAMP = 2 };  // imagewrap.rs

// sinc_1d (utils/mod.rs:207-217) is defined in camera.hpp (the Lanczos filter shares it).
inline int32_t modulo(int32_t a, int32_t b) {  // utils/mod.rs:219-223
  int32_t n = a / b;
  int32_t x = a - n * b;
  return x < 0 ? x + b : x;
}
inline int32_t iclamp(int32_t v, int32_t lo, int32_t hi) { return v < lo ? lo : (v > hi ? hi : v); }

struct ResampleWeight {
  int32_t first_texel;
  float weights[4];
};
inline std::vector<ResampleWeight> resample_weights(size_t oldres, size_t newres) {  // mipmap.rs:22-43
  std::vector<ResampleWeight> out(newres);
  const float filter_width = 2.0f;
  for (size_t i = 0; i < newres; ++i) {
    float center = ((float)i + 0.5f) * (float)oldres / (float)newres;
    int32_t first_texel = f2i(std::floor((center - filter_width) + 0.5f));
    ResampleWeight w;
    w.first_texel = first_texel;
    for (int j = 0; j < 4; ++j) {
      float pos = (float)(first_texel + j) + 0.5f;
      w.weights[j] = sinc_1d((pos - center) / filter_width, 2.0f);
    }
    float sum = 0.0f;  // iter().sum::<f32>()
    for (int j = 0; j < 4; ++j) sum = sum + w.weights[j];
    float inv = 1.0f / sum;
    for (int j = 0; j < 4; ++j) w.weights[j] *= inv;
    out[i] = w;
  }
  return out;
}

struct MipLevel {
  size_t w = 0, h = 0;
  std::vector<RGB> px;  // row-major (BlockedVec is a storage layout only, utils/blocked_vec.rs)
};

struct MIPMap {
  size_t width = 0, height = 0;
  std::vector<MipLevel> pyramid;
  bool do_trilinear = false;
  float max_anisotropy = 1.0f;
  int wrap = WRAP_REPEAT;

  static RGB texel_at(const MipLevel& l, int32_t s_, int32_t t_, int wm) {  // mipmap.rs:112-138
    int32_t s, t;
    if (wm == WRAP_REPEAT) {
      s = modulo(s_, (int32_t)l.w);
      t = modulo(t_, (int32_t)l.h);
    } else if (wm == WRAP_CLAMP) {
      s = iclamp(s_, 0, (int32_t)l.w - 1);
      t = iclamp(t_, 0, (int32_t)l.h - 1);
    } else {
      if (s_ < 0 || s_ >= (int32_t)l.w || t_ < 0 || t_ >= (int32_t)l.h) return RGB(0.f, 0.f, 0.f);
      s = s_;
      t = t_;
    }
    return l.px[(size_t)t * l.w + (size_t)s];
  }

  // mipmap.rs:45-106, including its in-place t pass (rows already resampled are read back).
  static void resize_pot(size_t w, size_t h, const std::vector<RGB>& pixels, int wm, size_t* wp, size_t* hp,
                         std::vector<RGB>* out) {
    size_t wpot = 1, hpot = 1;
    while (wpot < w) wpot <<= 1;
    while (hpot < h) hpot <<= 1;
    std::vector<RGB> np;
    np.reserve(wpot * hpot);
    auto get_orig = [&](const ResampleWeight& r, int j, size_t dim, int32_t* o) {
      int32_t ft = r.first_texel + j;
      int32_t orig = wm == WRAP_REPEAT ? modulo(ft, (int32_t)dim)
                                       : (wm == WRAP_CLAMP ? iclamp(ft, 0, (int32_t)dim - 1) : ft);
      if (orig >= 0 && orig < (int32_t)dim) {
        *o = orig;
        return true;
      }
      return false;
    };
    auto sw = resample_weights(w, wpot);
    for (size_t t = 0; t < h; ++t)
      for (size_t s = 0; s < wpot; ++s) {
        RGB acc(0.f, 0.f, 0.f);  // Sum for Spectrum: fold(zero, |acc, x| x + acc)
        for (int j = 0; j < 4; ++j) {
          int32_t o;
          if (get_orig(sw[s], j, w, &o)) acc = pixels[t * w + (size_t)o] * sw[s].weights[j] + acc;
        }
        np.push_back(acc);
      }
    for (size_t t = h; t < hpot; ++t)
      for (size_t s = 0; s < wpot; ++s) np.push_back(pixels[0]);
    auto tw = resample_weights(h, hpot);
    for (size_t s = 0; s < wpot; ++s)
      for (size_t t = 0; t < hpot; ++t) {
        RGB acc(0.f, 0.f, 0.f);
        for (int j = 0; j < 4; ++j) {
          int32_t o;
          if (get_orig(tw[t], j, h, &o)) acc = np[(size_t)o * wpot + s] * tw[t].weights[j] + acc;
        }
        np[t * wpot + s] = acc;
      }
    *wp = wpot;
    *hp = hpot;
    *out = std::move(np);
  }

  static size_t ulog2(size_t x) {  // mipmap.rs:108-110: bits - leading_zeros
    size_t n = 0;
    while (x) {
      ++n;
      x >>= 1;
    }
    return n;
  }

  MIPMap() {}
  MIPMap(size_t w, size_t h, const std::vector<RGB>& pixels, bool tri, float max_aniso, int wm) {  // :159-204
    std::vector<RGB> pot;
    auto is_pot = [](size_t v) { return v && !(v & (v - 1)); };
    if (!is_pot(w) || !is_pot(h))
      resize_pot(w, h, pixels, wm, &width, &height, &pot);
    else {
      width = w;
      height = h;
      pot = pixels;
    }
    MipLevel l0;
    l0.w = width;
    l0.h = height;
    l0.px = std::move(pot);
    pyramid.push_back(std::move(l0));
    size_t num_levels = ulog2(std::max(width, height));
    for (size_t i = 1; i < num_levels; ++i) {
      const MipLevel& last = pyramid.back();
      MipLevel nl;
      nl.w = std::max<size_t>(last.w / 2, 1);
      nl.h = std::max<size_t>(last.h / 2, 1);
      nl.px.resize(nl.w * nl.h);
      for (int32_t t = 0; t < (int32_t)nl.h; ++t)
        for (int32_t s = 0; s < (int32_t)nl.w; ++s) {
          RGB t0 = texel_at(last, 2 * s, 2 * t, wm), t1 = texel_at(last, 2 * s + 1, 2 * t, wm);
          RGB t2 = texel_at(last, 2 * s, 2 * t + 1, wm), t3 = texel_at(last, 2 * s + 1, 2 * t + 1, wm);
          nl.px[(size_t)t * nl.w + (size_t)s] = (((t0 + t1) + t2) + t3) * 0.25f;
        }
      pyramid.push_back(std::move(nl));
    }
    do_trilinear = tri;
    max_anisotropy = max_aniso;
    wrap = wm;
  }

  size_t levels() const { return pyramid.size(); }

  RGB triangle(size_t level_, float s_, float t_) const {  // :212-226
    size_t level = std::min(level_, levels() - 1);
    const MipLevel& l = pyramid[level];
    float s = s_ * (float)l.w - 0.5f, t = t_ * (float)l.h - 0.5f;
    int32_t s0 = f2i(std::floor(s)), t0 = f2i(std::floor(t));
    float ds = s - (float)s0, dt = t - (float)t0;
    return ((texel_at(l, s0, t0, wrap) * (1.0f - ds) * (1.0f - dt) + texel_at(l, s0, t0 + 1, wrap) * (1.0f - ds) * dt) +
            texel_at(l, s0 + 1, t0, wrap) * ds * (1.0f - dt)) +
           texel_at(l, s0 + 1, t0 + 1, wrap) * ds * dt;
  }

  static size_t f2usize(float v) {  // Rust `as usize`: saturating, NaN -> 0
    if (!(v > 0.0f)) return 0;
    if (v >= 1.8446744073709552e19f) return (size_t)-1;
    return (size_t)v;
  }

  RGB pyramid_lookup(float s, float t, float width_) const {  // :228-241
    float level = (float)levels() - 1.0f + std::log2(rmax(width_, 1e-8f));
    if (level < 0.0f) return triangle(0, s, t);
    if (level >= (float)(levels() - 1)) return texel_at(pyramid.back(), 0, 0, wrap);
    size_t ilevel = f2usize(level);
    float delta = level - (float)ilevel;
    RGB t0 = triangle(ilevel + 1, s, t);
    RGB t1 = triangle(ilevel, s, t);
    return t0 * (1.0f - delta) + t1 * delta;  // t0.lerp_with(t1, delta)
  }

  RGB ewa(size_t level, float s_, float t_, float ds0_, float dt0_, float ds1_, float dt1_) const {  // :243-300
    if (level >= levels()) return texel_at(pyramid.back(), 0, 0, wrap);
    const MipLevel& l = pyramid[level];
    float s = s_ * (float)l.w - 0.5f, t = t_ * (float)l.h - 0.5f;
    float ds0 = ds0_ * (float)l.w, dt0 = dt0_ * (float)l.h;
    float ds1 = ds1_ * (float)l.w, dt1 = dt1_ * (float)l.h;
    float a = dt0 * dt0 + dt1 * dt1 + 1.0f;
    float b = -2.0f * (ds0 * dt0 + ds1 * dt1);
    float c = ds0 * ds0 + ds1 * ds1 + 1.0f;
    float inv_f = 1.0f / (a * c - b * b * 0.25f);
    a = a * inv_f;
    b = b * inv_f;
    c = c * inv_f;
    float det = -b * b + 4.0f * a * c;
    float inv_det = 1.0f / det;
    float u_sqrt = std::sqrt(det * c), v_sqrt = std::sqrt(det * a);
    int32_t s0 = f2i(std::ceil(s - 2.0f * inv_det * u_sqrt)), s1 = f2i(std::floor(s + 2.0f * inv_det * u_sqrt));
    int32_t t0 = f2i(std::ceil(t - 2.0f * inv_det * v_sqrt)), t1 = f2i(std::floor(t + 2.0f * inv_det * v_sqrt));
    RGB sum(0.f, 0.f, 0.f);
    float sum_wts = 0.0f;
    const float INV_EXP_2 = 0.13533528323f;
    for (int64_t it = t0; it < (int64_t)t1 + 1; ++it) {
      float tt = (float)(int32_t)it - t;
      for (int64_t is = s0; is < (int64_t)s1 + 1; ++is) {
        float ss = (float)(int32_t)is - s;
        float r2 = a * ss * ss + b * ss * tt + c * tt * tt;
        if (r2 < 1.0f) {
          float weight = std::exp(-2.0f * r2) - INV_EXP_2;
          sum = sum + texel_at(l, (int32_t)is, (int32_t)it, wrap) * weight;
          sum_wts = sum_wts + weight;
        }
      }
    }
    return RGB(sum.c[0] / sum_wts, sum.c[1] / sum_wts, sum.c[2] / sum_wts);
  }

  RGB lookup(float s, float t, float dsdx, float dtdx, float dsdy, float dtdy) const {  // :302-341
    if (do_trilinear) {
      float width_ = rmax(rmax(rmax(std::fabs(dsdx), std::fabs(dtdx)), std::fabs(dsdy)), std::fabs(dtdy));
      return pyramid_lookup(s, t, 2.0f * width_);
    }
    float ds0, dt0, ds1, dt1;
    if (dsdx * dsdx + dtdx * dtdx > dsdy * dsdy + dtdy * dtdy) {
      ds0 = dsdx; dt0 = dtdx; ds1 = dsdy; dt1 = dtdy;
    } else {
      ds0 = dsdy; dt0 = dtdy; ds1 = dsdx; dt1 = dtdx;
    }
    float major_length = std::sqrt(ds0 * ds0 + dt0 * dt0);
    float minor_length = std::sqrt(ds1 * ds1 + dt1 * dt1);
    float max_major_length = minor_length * max_anisotropy;
    float sds1 = ds1, sdt1 = dt1, sminor = minor_length;
    if (max_major_length < major_length && minor_length > 0.0f) {
      float scale = major_length / (minor_length * max_anisotropy);
      sds1 = ds1 * scale;
      sdt1 = dt1 * scale;
      sminor = minor_length * scale;
    }
    if (sminor == 0.0f) return triangle(0, s, t);
    float lod = rmax((float)levels() - 1.0f + std::log2(minor_length), 0.0f);
    size_t ilod = f2usize(std::floor(lod));
    float d = lod - (float)ilod;
    RGB e0 = ewa(ilod + 0, s, t, ds0, dt0, sds1, sdt1);
    RGB e1 = ewa(ilod + 1, s, t, ds0, dt0, sds1, sdt1);
    return e0 * (1.0f - d) + e1 * d;
  }
};

// imagemap.rs:96-173: texels (read_image: byte / 255) -> (s * scale).powf(gamma) per channel
// (Spectrum cache) or (s.y() * scale).powf(gamma) (f32 cache); unreadable file -> 1x1 scale^gamma.
inline MIPMap make_image_mipmap(const float* rgb, size_t w, size_t h, bool spectrum, bool tri, float max_aniso, int wm,
                                float scale, float gamma) {
  std::vector<RGB> px;
  if (!rgb || w == 0 || h == 0) {
    float v = std::pow(scale, gamma);
    px.push_back(RGB(v, v, v));
    return MIPMap(1, 1, px, tri, max_aniso, wm);
  }
  px.resize(w * h);
  for (size_t i = 0; i < w * h; ++i) {
    float r = rgb[3 * i], g = rgb[3 * i + 1], b = rgb[3 * i + 2];
    if (spectrum) {
      px[i] = RGB(std::pow(r * scale, gamma), std::pow(g * scale, gamma), std::pow(b * scale, gamma));
    } else {
      float y = 0.212671f * r + 0.715160f * g + 0.072169f * b;  // Spectrum::y -> rgb_to_xyz[1], spectrum.rs:37-41
      float v = std::pow(y * scale, gamma);
      px[i] = RGB(v, v, v);
    }
  }
  return MIPMap(w, h, px, tri, max_aniso, wm);
}

}  // namespace orc
