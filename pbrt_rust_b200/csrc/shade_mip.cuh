// ImageTexture lookups of the shading stage (sm_100a): MIPMap::{texel_at, triangle, pyramid_lookup,
// ewa, lookup} (src/texture/mipmap.rs:112-341), operation order as written.  Reads texels through
// __ldg only, so the same source compiles as host code (PB_HOST_CHECK, tests/devsrc/) and the CPU
// test-suite can run it against the oracle; the product runs it on the GPU.
#pragma once
#include "shade_tex.cuh"  // f3, PB_NOINLINE, the mip_lookup declaration

// ---- ImageTexture: MIPMap::lookup (texture/mipmap.rs:206-341) -----------------------------------
// Level l of a map starts where level l-1 ends (row-major RGB float4 texels, include/pbrtb200.h).
struct DMipLevel {
  const float4* px;
  int w, h;
};
PB_DEV DMipLevel mip_level(const float4* __restrict__ texels, const pbrtb200_mipmap& mm, uint32_t level) {
  uint64_t off = mm.texel_offset;
  uint32_t w = mm.width, h = mm.height;
  for (uint32_t l = 0; l < level; ++l) {
    off += (uint64_t)w * h;
    w = w > 1 ? w >> 1 : 1;
    h = h > 1 ? h >> 1 : 1;
  }
  DMipLevel r;
  r.px = texels + off;
  r.w = (int)w;
  r.h = (int)h;
  return r;
}
PB_DEV int mip_modulo(int a, int b) {  // utils/mod.rs:219-223
  const int x = a - (a / b) * b;
  return x < 0 ? x + b : x;
}
PB_DEV f3 mip_texel(const DMipLevel& l, int s_, int t_, uint32_t wrap) {  // texel_at, mipmap.rs:112-138
  int s, t;
  if (wrap == PBRTB200_WRAP_REPEAT) {
    s = mip_modulo(s_, l.w);
    t = mip_modulo(t_, l.h);
  } else if (wrap == PBRTB200_WRAP_CLAMP) {
    s = min(max(s_, 0), l.w - 1);
    t = min(max(t_, 0), l.h - 1);
  } else {
    if (s_ < 0 || s_ >= l.w || t_ < 0 || t_ >= l.h) return mk3(0.f, 0.f, 0.f);
    s = s_;
    t = t_;
  }
  const float4 v = __ldg(l.px + (size_t)t * (size_t)l.w + (size_t)s);
  return mk3(v.x, v.y, v.z);
}
PB_DEV int iadd_wrap(int a, int b) { return (int)((uint32_t)a + (uint32_t)b); }  // release-mode i32 add
PB_DEV f3 mip_triangle(const float4* __restrict__ texels, const pbrtb200_mipmap& mm, uint32_t level_, float s_, float t_) {  // :212-226
  const uint32_t level = level_ < mm.n_levels - 1 ? level_ : mm.n_levels - 1;
  const DMipLevel l = mip_level(texels, mm, level);
  const float s = s_ * (float)l.w - 0.5f, t = t_ * (float)l.h - 0.5f;
  const int s0 = f2i_sat(floorf(s)), t0 = f2i_sat(floorf(t));
  const float ds = s - (float)s0, dt = t - (float)t0;
  const int s1 = iadd_wrap(s0, 1), t1 = iadd_wrap(t0, 1);
  return ((mip_texel(l, s0, t0, mm.wrap) * (1.0f - ds) * (1.0f - dt) + mip_texel(l, s0, t1, mm.wrap) * (1.0f - ds) * dt) +
          mip_texel(l, s1, t0, mm.wrap) * ds * (1.0f - dt)) +
         mip_texel(l, s1, t1, mm.wrap) * ds * dt;
}
PB_DEV uint32_t f2level(float v) {  // `as usize` (saturating, NaN -> 0); levels never exceed 32
  if (!(v > 0.0f)) return 0u;
  return v >= 4.0e9f ? 0xFFFFFFF0u : (uint32_t)v;
}
PB_DEV f3 mip_ewa(const float4* __restrict__ texels, const pbrtb200_mipmap& mm, uint32_t level, float s_, float t_, float ds0_,
                  float dt0_, float ds1_, float dt1_) {  // :243-300
  if (level >= mm.n_levels) return mip_texel(mip_level(texels, mm, mm.n_levels - 1), 0, 0, mm.wrap);
  const DMipLevel l = mip_level(texels, mm, level);
  const float s = s_ * (float)l.w - 0.5f, t = t_ * (float)l.h - 0.5f;
  const float ds0 = ds0_ * (float)l.w, dt0 = dt0_ * (float)l.h;
  const float ds1 = ds1_ * (float)l.w, dt1 = dt1_ * (float)l.h;
  float a = dt0 * dt0 + dt1 * dt1 + 1.0f;
  float b = -2.0f * (ds0 * dt0 + ds1 * dt1);
  float c = ds0 * ds0 + ds1 * ds1 + 1.0f;
  const float inv_f = 1.0f / (a * c - b * b * 0.25f);
  a = a * inv_f;
  b = b * inv_f;
  c = c * inv_f;
  const float det = -b * b + 4.0f * a * c;
  const float inv_det = 1.0f / det;
  const float u_sqrt = sqrtf(det * c), v_sqrt = sqrtf(det * a);
  const int s0 = f2i_sat(ceilf(s - 2.0f * inv_det * u_sqrt)), s1 = f2i_sat(floorf(s + 2.0f * inv_det * u_sqrt));
  const int t0 = f2i_sat(ceilf(t - 2.0f * inv_det * v_sqrt)), t1 = f2i_sat(floorf(t + 2.0f * inv_det * v_sqrt));
  f3 sum = mk3(0.f, 0.f, 0.f);
  float sum_wts = 0.0f;
  // The reference walks the ellipse's whole bounding box (mipmap.rs:281-296) and adds the texels with
  // r2 < 1 in row-major order.  As written it picks the level from the UNCLAMPED minor axis
  // (mipmap.rs:335), so at grazing angles the box holds 10^6 .. 10^8 texels.  Texels outside the
  // ellipse add nothing, so each row is cut to the interval where r2 < 1 can hold: the roots of
  // a ss^2 + (b tt) ss + (c tt^2 - 1) = 0, evaluated in double and widened by two texels plus 1e-5 of
  // their magnitude — far more than the band in which float rounding of r2 can flip the comparison
  // (|r2 - 1| < ~1e-6 spans |ss| * 1e-6 texels).  The exact float test still decides every texel
  // inside the interval, so the accepted set, the order and hence the sums are unchanged bit for bit;
  // the work drops from the box to the ellipse (a rotated 8:1 ellipse fills ~1/6 of its box).
  const bool cut = a > 0.0f && a < PB_F32_MAX && b == b && c == c && fabsf(b) < PB_F32_MAX && fabsf(c) < PB_F32_MAX;
  for (long long it = t0; it <= (long long)t1; ++it) {
    const float tt = (float)(int)it - t;
    long long is0 = s0, is1 = s1;
    if (cut) {
      const double bt = (double)b * (double)tt, ct = (double)c * (double)tt * (double)tt - 1.0;
      const double disc = bt * bt - 4.0 * (double)a * ct;
      if (disc < -1e-9 * (bt * bt + 4.0 * fabs((double)a * ct))) continue;  // the row misses the ellipse
      const double half = sqrt(disc > 0.0 ? disc : 0.0) / (2.0 * (double)a), mid = -bt / (2.0 * (double)a);
      const double lo = (double)s + mid - half, hi = (double)s + mid + half;
      const double m = 2.0 + 1e-5 * (fabs(lo) + fabs(hi) + fabs((double)s));
      const double lo2 = floor(lo - m), hi2 = ceil(hi + m);
      if (lo2 > (double)is0) is0 = lo2 < 9.2e18 ? (long long)lo2 : is1 + 1;
      if (hi2 < (double)is1) is1 = hi2 > -9.2e18 ? (long long)hi2 : is0 - 1;
    }
    for (long long is = is0; is <= is1; ++is) {
      const float ss = (float)(int)is - s;
      const float r2 = a * ss * ss + b * ss * tt + c * tt * tt;
      if (r2 < 1.0f) {
        const float weight = expf(-2.0f * r2) - 0.13533528323f;
        sum = sum + mip_texel(l, (int)is, (int)it, mm.wrap) * weight;
        sum_wts = sum_wts + weight;
      }
    }
  }
  return mk3(sum.x / sum_wts, sum.y / sum_wts, sum.z / sum_wts);
}
// Not inlined: the lookup is a cold, register-hungry path next to constant/checker textures, and
// inlining it at every tex_eval site pushed k_shade's spill stack from 120 to 216 bytes.
PB_NOINLINE f3 mip_lookup(const float4* __restrict__ texels, const pbrtb200_mipmap* __restrict__ mmp,
                                      float s, float t, float dsdx, float dtdx, float dsdy, float dtdy) {  // :302-341
  const pbrtb200_mipmap mm = *mmp;
  if (mm.do_trilinear) {
    const float width = fmaxf(fmaxf(fmaxf(fabsf(dsdx), fabsf(dtdx)), fabsf(dsdy)), fabsf(dtdy));
    // pyramid_lookup (:228-241)
    const float level = (float)mm.n_levels - 1.0f + log2f(fmaxf(2.0f * width, 1e-8f));
    if (level < 0.0f) return mip_triangle(texels, mm, 0, s, t);
    if (level >= (float)(mm.n_levels - 1)) return mip_texel(mip_level(texels, mm, mm.n_levels - 1), 0, 0, mm.wrap);
    const uint32_t ilevel = f2level(level);
    const float delta = level - (float)ilevel;
    const f3 t0 = mip_triangle(texels, mm, ilevel + 1, s, t);
    const f3 t1 = mip_triangle(texels, mm, ilevel, s, t);
    return t0 * (1.0f - delta) + t1 * delta;  // t0.lerp_with(t1, delta), utils/mod.rs:20
  }
  float ds0, dt0, ds1, dt1;
  if (dsdx * dsdx + dtdx * dtdx > dsdy * dsdy + dtdy * dtdy) {
    ds0 = dsdx; dt0 = dtdx; ds1 = dsdy; dt1 = dtdy;
  } else {
    ds0 = dsdy; dt0 = dtdy; ds1 = dsdx; dt1 = dtdx;
  }
  const float major_length = sqrtf(ds0 * ds0 + dt0 * dt0);
  const float minor_length = sqrtf(ds1 * ds1 + dt1 * dt1);
  const float max_major_length = minor_length * mm.max_anisotropy;
  float sds1 = ds1, sdt1 = dt1, sminor = minor_length;
  if (max_major_length < major_length && minor_length > 0.0f) {
    const float scale = major_length / (minor_length * mm.max_anisotropy);
    sds1 = ds1 * scale;
    sdt1 = dt1 * scale;
    sminor = minor_length * scale;
  }
  if (sminor == 0.0f) return mip_triangle(texels, mm, 0, s, t);
  const float lod = fmaxf((float)mm.n_levels - 1.0f + log2f(minor_length), 0.0f);
  const uint32_t ilod = f2level(floorf(lod));
  const float d = lod - (float)ilod;
  const f3 e0 = mip_ewa(texels, mm, ilod, s, t, ds0, dt0, sds1, sdt1);
  const f3 e1 = mip_ewa(texels, mm, ilod + 1, s, t, ds0, dt0, sds1, sdt1);
  return e0 * (1.0f - d) + e1 * d;
}

