// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// Surface-interaction construction, textures, matte/plastic BSDFs, lights and the Whitted light
// loop of pbrt_rust, restated line-faithfully, with the SURVEY §0.2 decisions (D2, D3, D7, D8, D9,
// D10, D11, D17).  Citations are relative to the reference root.
#pragma once
#include "accel.hpp"
#include "camera.hpp"

namespace orc {

// spectrum.rs — RGB variant only (Spectrum::from(f32) -> RGB, spectrum.rs:495-499).
struct RGB {
  float c[3] = {0.f, 0.f, 0.f};
  RGB() = default;
  explicit RGB(float f) { c[0] = c[1] = c[2] = f; }
  RGB(float r, float g, float b) {
    c[0] = r;
    c[1] = g;
    c[2] = b;
  }
  bool is_black() const { return c[0] == 0.0f && c[1] == 0.0f && c[2] == 0.0f; }  // :425
  bool has_nans() const { return std::isnan(c[0]) || std::isnan(c[1]) || std::isnan(c[2]); }
  RGB clamp(float a, float b) const {
    return RGB(rclamp(c[0], a, b), rclamp(c[1], a, b), rclamp(c[2], a, b));
  }
};
inline RGB operator+(const RGB& a, const RGB& b) {
  return RGB(a.c[0] + b.c[0], a.c[1] + b.c[1], a.c[2] + b.c[2]);
}
inline RGB operator-(const RGB& a, const RGB& b) {
  return RGB(a.c[0] - b.c[0], a.c[1] - b.c[1], a.c[2] - b.c[2]);
}
inline RGB operator*(const RGB& a, const RGB& b) {
  return RGB(a.c[0] * b.c[0], a.c[1] * b.c[1], a.c[2] * b.c[2]);
}
inline RGB operator/(const RGB& a, const RGB& b) {
  return RGB(a.c[0] / b.c[0], a.c[1] / b.c[1], a.c[2] / b.c[2]);
}
inline RGB operator*(const RGB& a, float s) { return RGB(a.c[0] * s, a.c[1] * s, a.c[2] * s); }
inline RGB operator*(float s, const RGB& a) { return a * s; }
inline RGB operator/(const RGB& a, float s) { return RGB(a.c[0] / s, a.c[1] / s, a.c[2] / s); }

// ---------------------------------------------------------------------------------------------
// Textures (texture/mod.rs:52-66, checkerboard.rs:24-95, uv.rs:20-26, mapping2d.rs:49-77,175-210)
struct Mapping2D {
  // 0 = UVMapping2D(su,sv,du,dv) ; 1 = PlanarMapping2D(vs,vt,ds,dt) ;
  // 2 = SphericalMapping2D(world_to_texture) ; 3 = CylindricalMapping2D(world_to_texture) ;
  // 4 = IdentityMapping3D(world_to_texture) (mapping3d.rs:43-66; used through map3)
  int kind = 0;
  float su = 1.f, sv = 1.f, du = 0.f, dv = 0.f;
  V3 vs{1, 0, 0}, vt{0, 1, 0};
  Transform w2t = Transform::translate(V3(0.f, 0.f, 0.f));
  // mapping2d.rs:120-133 / 155-166: the point as a unit vector ((1,0,0) at the origin)
  V3 unit_vec(const V3& p) const {
    V3 v = w2t.pt(p);
    if (v.x == 0.0f && v.y == 0.0f && v.z == 0.0f) return V3(1.0f, 0.0f, 0.0f);
    return normalize(v);
  }
  void circular(const V3& p, float* s, float* t) const {
    V3 vec = unit_vec(p);
    if (kind == 2) {  // sphere(): spherical_theta / spherical_phi (vector.rs:210-217)
      const float FRAC_1_PI = 0.318309886183790671537767526745028724f;
      float theta = std::acos(rclamp(vec.z, -1.0f, 1.0f));
      float phi = std::atan2(vec.y, vec.x);
      if (phi < 0.0f) phi = phi + 2.0f * PI_F;
      *s = theta * FRAC_1_PI;
      *t = phi * FRAC_1_PI * 0.5f;
    } else {  // cylinder()
      *s = (PI_F + std::atan2(vec.y, vec.x)) / (2.0f * PI_F);
      *t = vec.z;
    }
  }
  void map(const DiffGeom& dg, float o[6]) const {  // (s, t, dsdx, dtdx, dsdy, dtdy)
    if (kind == 0) {
      o[0] = su * dg.u + du;
      o[1] = sv * dg.v + dv;
      o[2] = su * dg.dudx;
      o[3] = sv * dg.dvdx;
      o[4] = su * dg.dudy;
      o[5] = sv * dg.dvdy;
    } else if (kind == 1) {
      V3 vec = dg.p;
      o[0] = du + dot(vec, vs);
      o[1] = dv + dot(vec, vt);
      o[2] = dot(vs, dg.dpdx);
      o[3] = dot(vt, dg.dpdx);
      o[4] = dot(vs, dg.dpdy);
      o[5] = dot(vt, dg.dpdy);
    } else {  // get_circular_differentials (mapping2d.rs:78-104)
      float s, t;
      circular(dg.p, &s, &t);
      const float delta = 0.1f;
      auto deal_with_singularity = [](float res) {
        if (res > 0.5f) return 1.0f - res;
        if (res < -0.5f) return -(res + 1.0f);
        return res;
      };
      float sx, tx, sy, ty;
      circular(dg.p + delta * dg.dpdx, &sx, &tx);
      circular(dg.p + delta * dg.dpdy, &sy, &ty);
      o[0] = s;
      o[1] = t;
      o[2] = (sx - s) / delta;
      o[3] = deal_with_singularity((tx - t) / delta);
      o[4] = (sy - s) / delta;
      o[5] = deal_with_singularity((ty - t) / delta);
    }
  }
  // IdentityMapping3D::map_dg (mapping3d.rs:57-63)
  void map3(const DiffGeom& dg, V3* p, V3* dpdx, V3* dpdy) const {
    *dpdx = w2t.vec(dg.dpdx);
    *dpdy = w2t.vec(dg.dpdy);
    *p = w2t.pt(dg.p);
  }
};

// texture/noise.rs ---------------------------------------------------------------------------
namespace noise_detail {
static const uint8_t PERM[256] = {  // NOISE_PERM (noise.rs:6-55; the table is stored twice there)
    151, 160, 137, 91,  90,  15,  131, 13,  201, 95,  96,  53,  194, 233, 7,   225, 140, 36,  103, 30,
    69,  142, 8,   99,  37,  240, 21,  10,  23,  190, 6,   148, 247, 120, 234, 75,  0,   26,  197, 62,
    94,  252, 219, 203, 117, 35,  11,  32,  57,  177, 33,  88,  237, 149, 56,  87,  174, 20,  125, 136,
    171, 168, 68,  175, 74,  165, 71,  134, 139, 48,  27,  166, 77,  146, 158, 231, 83,  111, 229, 122,
    60,  211, 133, 230, 220, 105, 92,  41,  55,  46,  245, 40,  244, 102, 143, 54,  65,  25,  63,  161,
    1,   216, 80,  73,  209, 76,  132, 187, 208, 89,  18,  169, 200, 196, 135, 130, 116, 188, 159, 86,
    164, 100, 109, 198, 173, 186, 3,   64,  52,  217, 226, 250, 124, 123, 5,   202, 38,  147, 118, 126,
    255, 82,  85,  212, 207, 206, 59,  227, 47,  16,  58,  17,  182, 189, 28,  42,  223, 183, 170, 213,
    119, 248, 152, 2,   44,  154, 163, 70,  221, 153, 101, 155, 167, 43,  172, 9,   129, 22,  39,  253,
    19,  98,  108, 110, 79,  113, 224, 232, 178, 185, 112, 104, 218, 246, 97,  228, 251, 34,  242, 193,
    238, 210, 144, 12,  191, 179, 162, 241, 81,  51,  145, 235, 249, 14,  239, 107, 49,  192, 214, 31,
    181, 199, 106, 157, 184, 84,  204, 176, 115, 121, 50,  45,  127, 4,   150, 254, 138, 236, 205, 93,
    222, 114, 67,  29,  24,  72,  243, 141, 128, 195, 78,  66,  215, 61,  156, 180};
inline uint32_t perm(uint32_t i) { return PERM[i & 255u]; }  // index < 512 into the doubled table
inline float grad(uint32_t x, uint32_t y, uint32_t z, float dx, float dy, float dz) {  // :57-62
  uint32_t h = perm(perm(perm(x) + y) + z) & 15u;
  float u = (h < 8 || h == 12 || h == 13) ? dx : dy;
  float v = (h < 4 || h == 12 || h == 13) ? dy : dz;
  return ((h & 1) == 0 ? u : -u) + ((h & 2) == 0 ? v : -v);
}
inline float noise_weight(float t) {  // :64-68
  float t3 = t * t * t;
  float t4 = t3 * t;
  return 6.0f * t4 * t - 15.0f * t4 + 10.0f * t3;
}
inline float lerp_with(float a, float b, float t) { return a * (1.0f - t) + b * t; }  // utils/mod.rs:20
}  // namespace noise_detail
inline float noise(float x, float y, float z) {  // noise.rs:70-103
  using namespace noise_detail;
  uint32_t ix = (uint32_t)(f2i(std::floor(x)) & 255);
  uint32_t iy = (uint32_t)(f2i(std::floor(y)) & 255);
  uint32_t iz = (uint32_t)(f2i(std::floor(z)) & 255);
  float dx = x - std::floor(x), dy = y - std::floor(y), dz = z - std::floor(z);
  float w000 = grad(ix, iy, iz, dx, dy, dz);
  float w100 = grad(ix + 1, iy, iz, dx - 1.0f, dy, dz);
  float w010 = grad(ix, iy + 1, iz, dx, dy - 1.0f, dz);
  float w110 = grad(ix + 1, iy + 1, iz, dx - 1.0f, dy - 1.0f, dz);
  float w001 = grad(ix, iy, iz + 1, dx, dy, dz - 1.0f);
  float w101 = grad(ix + 1, iy, iz + 1, dx - 1.0f, dy, dz - 1.0f);
  float w011 = grad(ix, iy + 1, iz + 1, dx, dy - 1.0f, dz - 1.0f);
  float w111 = grad(ix + 1, iy + 1, iz + 1, dx - 1.0f, dy - 1.0f, dz - 1.0f);
  float wx = noise_weight(dx), wy = noise_weight(dy), wz = noise_weight(dz);
  float x00 = lerp_with(w000, w100, wx);
  float x10 = lerp_with(w010, w110, wx);
  float x01 = lerp_with(w001, w101, wx);
  float x11 = lerp_with(w011, w111, wx);
  float y0 = lerp_with(x00, x10, wy);
  float y1 = lerp_with(x01, x11, wy);
  return lerp_with(y0, y1, wz);
}
inline float smoothstep(float mn, float mx, float value) {  // noise.rs:107-110
  float v = rclamp((value - mn) / (mx - mn), 0.0f, 1.0f);
  return v * v * (-2.0f * v + 3.0f);
}
// fbm (noise.rs:112-127) and turbulence (:129-145).  As written, turbulence takes |noise| only
// for the partial octave (pbrt-v2 takes it in the loop as well).
inline float fbm_or_turbulence(bool turb, const V3& p, const V3& dpdx, const V3& dpdy, float omega,
                               int max_octaves) {
  float s2 = rmax(length_squared(dpdx), length_squared(dpdy));
  float foctaves = rclamp(-1.0f - 0.5f * std::log2(s2), 0.0f, (float)max_octaves);
  int32_t octaves = f2i(std::floor(foctaves));
  float sum = 0.0f, lambda = 1.0f, o = 1.0f;
  for (int32_t i = 0; i < octaves; ++i) {
    float v = noise(lambda * p.x, lambda * p.y, lambda * p.z);
    sum = sum + o * v;
    lambda = lambda * 1.99f;
    o = o * omega;
  }
  float partial_octave = foctaves - std::floor(foctaves);
  float n = noise(lambda * p.x, lambda * p.y, lambda * p.z);
  if (turb) n = std::fabs(n);
  return sum + o * smoothstep(0.3f, 0.7f, partial_octave) * n;
}
}  // namespace orc
#include "mipmap.hpp"  // MIPMap over RGB texels (needs RGB above)
namespace orc {
struct Texture {
  // 0 = Constant, 1 = Checkerboard2D, 2 = UV, 3 = ImageTexture (tex1 = mipmap index),
  // 4 = Scale (tex1 * tex2), 5 = Mix (tex1.lerp(tex2, tex3)), 6 = Bilerp (v00, v01, v10, v11),
  // 7 = Dots (tex1 inside, tex2 outside), 8 = FBm, 9 = Wrinkled (value.c[0] = omega, aa = octaves)
  int kind = 0;
  RGB value;     // Constant (float textures use c[0])
  RGB bil[4];    // Bilerp corner values v00, v01, v10, v11
  Mapping2D mapping;
  int tex1 = 0, tex2 = 0, tex3 = 0;  // children (indices)
  int aa = 0;                         // Checkerboard: 0 = NONE, 1 = CLOSEDFORM ; FBm/Wrinkled: octaves
};
struct TextureTable {
  std::vector<Texture> t;
  std::vector<MIPMap> mips;
  RGB eval(int id, const DiffGeom& dg) const {
    const Texture& tx = t[(size_t)id];
    switch (tx.kind) {
      case 0:
        return tx.value;
      case 3: {  // imagemap.rs:200-206
        float m[6];
        tx.mapping.map(dg, m);
        return mips[(size_t)tx.tex1].lookup(m[0], m[1], m[2], m[3], m[4], m[5]);
      }
      case 2: {
        float m[6];
        tx.mapping.map(dg, m);
        return RGB(m[0] - std::floor(m[0]), m[1] - std::floor(m[1]), 0.0f);
      }
      case 4:  // ScaleTexture (texture/mod.rs:81-85)
        return eval(tx.tex1, dg) * eval(tx.tex2, dg);
      case 5: {  // MixTexture (mix.rs:22-27): tex1.lerp(tex2, amount)
        RGB a = eval(tx.tex1, dg), b = eval(tx.tex2, dg);
        float amt = eval(tx.tex3, dg).c[0];
        return a * (1.0f - amt) + b * amt;
      }
      case 6: {  // BilerpTexture (bilerp.rs:28-35)
        float m[6];
        tx.mapping.map(dg, m);
        RGB tmp1 = tx.bil[0] * (1.0f - m[0]) + tx.bil[2] * m[0];
        RGB tmp2 = tx.bil[1] * (1.0f - m[0]) + tx.bil[3] * m[0];
        return tmp1 * (1.0f - m[1]) + tmp2 * m[1];
      }
      case 7: {  // DotsTexture (dots.rs:23-46)
        float m[6];
        tx.mapping.map(dg, m);
        float s = m[0], t_ = m[1];
        float s_cell = std::floor(s + 0.5f), t_cell = std::floor(t_ + 0.5f);
        if (noise(s_cell + 0.5f, t_cell + 0.5f, 0.5f) > 0.0f) {
          const float radius = 0.35f;
          const float max_shift = 0.5f - radius;
          float s_center = s_cell + max_shift * noise(s_cell + 1.5f, t_cell + 2.8f, 0.5f);
          float t_center = t_cell + max_shift * noise(s_cell + 4.5f, t_cell + 9.8f, 0.5f);
          float ds = s - s_center, dt = t_ - t_center;
          if (ds * ds + dt * dt < radius * radius) return eval(tx.tex1, dg);
          return eval(tx.tex2, dg);
        }
        return eval(tx.tex2, dg);
      }
      case 8:
      case 9: {  // FBmTexture / WrinkledTexture (fbm.rs:21-26, 42-47)
        V3 p, dpdx, dpdy;
        tx.mapping.map3(dg, &p, &dpdx, &dpdy);
        return RGB(fbm_or_turbulence(tx.kind == 9, p, dpdx, dpdy, tx.value.c[0], tx.aa));
      }
      default: {
        float m[6];
        tx.mapping.map(dg, m);
        float s = m[0], t_ = m[1];
        auto point_sample = [&]() {
          int32_t a = f2i(std::floor(s)), b = f2i(std::floor(t_));
          int32_t sum = (int32_t)((uint32_t)a + (uint32_t)b);
          return (sum % 2 == 0) ? eval(tx.tex1, dg) : eval(tx.tex2, dg);
        };
        if (tx.aa == 0) return point_sample();
        float ds = rmax(std::fabs(m[2]), std::fabs(m[4]));
        float dt = rmax(std::fabs(m[3]), std::fabs(m[5]));
        float s0 = s - ds, t0 = t_ - dt, s1 = s + ds, t1 = t_ + dt;
        if (std::floor(s0) == std::floor(s1) && std::floor(t0) == std::floor(t1))
          return point_sample();
        auto bump_int = [](float x) {
          float half_x = x / 2.0f;
          return std::floor(half_x) + 2.0f * rmax(half_x - std::floor(half_x) - 0.5f, 0.0f);
        };
        float sint = ds > 0.0f ? (bump_int(s1) - bump_int(s0)) / (2.0f * ds) : 0.0f;
        float tint = dt > 0.0f ? (bump_int(t1) - bump_int(t0)) / (2.0f * dt) : 0.0f;
        float area_sq = (ds > 1.0f || dt > 1.0f) ? 0.5f : sint + tint - 2.0f * sint * tint;
        RGB a = eval(tx.tex1, dg), b = eval(tx.tex2, dg);
        return a * (1.0f - area_sq) + b * area_sq;  // Lerp::lerp_with
      }
    }
  }
};

struct Material {
  int kind = 0;  // 0 = Matte(kd, sigma), 1 = Plastic(kd, ks, roughness)
  int kd = 0, sigma = 0, ks = 0, roughness = 0;  // texture ids
  int bump = -1;  // bump_map: Option<..> (matte.rs:15, plastic.rs:18): displacement texture or -1
};

// material::bump (material/mod.rs:23-77)
inline DiffGeom bump_dg(const TextureTable& tt, int d, const DiffGeom& dg_geom, const DiffGeom& dg_shading) {
  DiffGeom dg_eval = dg_shading;
  float du = 0.5f * (std::fabs(dg_shading.dudx) + std::fabs(dg_shading.dudy));
  if (du == 0.0f) du = 0.1f;
  dg_eval.p = dg_shading.p + du * dg_shading.dpdu;
  dg_eval.u = dg_shading.u + du;
  dg_eval.nn = normalize(cross(dg_shading.dpdu, dg_shading.dpdv) + du * dg_shading.dndu);
  float u_displace = tt.eval(d, dg_eval).c[0];
  float dv = 0.5f * (std::fabs(dg_shading.dvdx) + std::fabs(dg_shading.dvdy));
  if (dv == 0.0f) dv = 0.1f;
  dg_eval.p = dg_shading.p + dv * dg_shading.dpdv;
  dg_eval.u = dg_shading.u;
  dg_eval.v = dg_shading.v + dv;
  dg_eval.nn = normalize(cross(dg_shading.dpdu, dg_shading.dpdv) + dv * dg_shading.dndv);
  float v_displace = tt.eval(d, dg_eval).c[0];
  float displace = tt.eval(d, dg_shading).c[0];
  DiffGeom dg_bump = dg_shading;
  dg_bump.dpdu = dg_shading.dpdu + (u_displace - displace) / du * dg_shading.nn + displace * dg_shading.dndu;
  dg_bump.dpdv = dg_shading.dpdv + (v_displace - displace) / dv * dg_shading.nn + displace * dg_shading.dndv;
  dg_bump.nn = normalize(cross(dg_bump.dpdu, dg_bump.dpdv));
  if (dg_shading.flip) dg_bump.nn = V3(-dg_bump.nn.x, -dg_bump.nn.y, -dg_bump.nn.z);
  dg_bump.nn = face_forward(dg_bump.nn, dg_geom.nn);
  return dg_bump;
}

// ---------------------------------------------------------------------------------------------
// bsdf/*
namespace bx {
inline float abs_cos_theta(const V3& v) { return std::fabs(v.z); }
inline float sin_theta2(const V3& v) { return rmax(0.f, 1.0f - v.z * v.z); }
inline float sin_theta(const V3& v) { return std::sqrt(sin_theta2(v)); }
inline float cos_phi(const V3& v) {
  float st = sin_theta(v);
  return st == 0.0f ? 1.0f : rclamp(v.x / st, -1.0f, 1.0f);
}
inline float sin_phi(const V3& v) {
  float st = sin_theta(v);
  return st == 0.0f ? 0.0f : rclamp(v.y / st, -1.0f, 1.0f);
}
}  // namespace bx

enum BxDFType : uint32_t {
  BSDF_REFLECTION = 1, BSDF_TRANSMISSION = 2, BSDF_DIFFUSE = 4, BSDF_GLOSSY = 8, BSDF_SPECULAR = 16,
  BSDF_ALL = 31
};

struct BxDF {
  int kind = 0;  // 0 Lambertian, 1 OrenNayar, 2 Microfacet(Blinn, dielectric 1.5/1.0)
  RGB r;
  float a = 0.f, b = 0.f;  // OrenNayar A,B ; Microfacet: a = blinn exponent
  uint32_t type() const {
    return kind == 2 ? (BSDF_REFLECTION | BSDF_GLOSSY) : (BSDF_REFLECTION | BSDF_DIFFUSE);
  }
  // fresnel.rs:62-91 (Dielectric arm), all three channels equal
  static float fresnel_dielectric(float cosi, float eta_i, float eta_t) {
    float ci = rclamp(cosi, -1.0f, 1.0f);
    float ei = eta_i, et = eta_t;
    if (cosi <= 0.0f) std::swap(ei, et);
    float sint = (ei / et) * std::sqrt(rmax(1.0f - ci * ci, 0.0f));
    if (sint >= 1.0f) return 1.0f;
    float cost = std::sqrt(rmax(1.0f - sint * sint, 0.0f));
    float aci = std::fabs(ci);
    float rparl = ((et * aci) - (ei * cost)) / ((et * aci) + (ei * cost));
    float rperp = ((ei * aci) - (et * cost)) / ((ei * aci) + (et * cost));
    return (rparl * rparl + rperp * rperp) / 2.0f;
  }
  RGB f(const V3& wo, const V3& wi) const {
    using namespace bx;
    if (kind == 0) {  // lambertian.rs:20-23
      float invpi = 1.0f / PI_F;
      return r * invpi;
    }
    if (kind == 1) {  // orennayar.rs:36-60
      float sinthetai = sin_theta(wi), sinthetao = sin_theta(wo);
      float maxcos = 0.0f;
      if (!(sinthetai < 1e-4f || sinthetao < 1e-4f)) {
        float sinphii = sin_phi(wi), cosphii = cos_phi(wi);
        float sinphio = sin_phi(wo), cosphio = cos_phi(wo);
        maxcos = rmax(cosphii * cosphio + sinphii * sinphio, 0.0f);
      }
      float sinalpha, tanbeta;
      if (abs_cos_theta(wi) > abs_cos_theta(wo)) {
        sinalpha = sinthetao;
        tanbeta = sinthetai / abs_cos_theta(wi);
      } else {
        sinalpha = sinthetai;
        tanbeta = sinthetao / abs_cos_theta(wo);
      }
      float invpi = 1.0f / PI_F;
      return r * invpi * (a + b * maxcos * sinalpha * tanbeta);
    }
    // microfacet.rs:86-99
    float cos_o = abs_cos_theta(wo), cos_i = abs_cos_theta(wi);
    if (cos_o == 0.0f || cos_i == 0.0f) return RGB(0.0f);
    V3 wh = normalize(wo + wi);
    float cos_h = dot(wi, wh);
    float F = fresnel_dielectric(cos_h, 1.5f, 1.0f);
    float invtwopi = 1.0f / (2.0f * PI_F);
    float D = (a + 2.0f) * invtwopi * std::pow(abs_cos_theta(wh), a);  // microfacet.rs:32-38
    float ndotwh = abs_cos_theta(wh), ndotwo = abs_cos_theta(wo), ndotwi = abs_cos_theta(wi);
    float wodotwh = abs_dot(wo, wh);
    float G = rmin(rmin(2.0f * ndotwh * ndotwo / wodotwh, 2.0f * ndotwh * ndotwi / wodotwh), 1.0f);
    return (r * D * G * RGB(F)) / (4.0f * cos_i * cos_o);
  }
};

struct BSDF {  // bsdf/mod.rs:58-149
  DiffGeom dg_shading;
  V3 nn, ng, sn, tn;
  BxDF bxdfs[2];
  int n_bxdfs = 0;
  BSDF(const DiffGeom& dgs, const V3& n_geom) : dg_shading(dgs) {  // :70-86
    nn = dgs.nn;
    tn = normalize(dgs.dpdu);
    sn = cross(nn, tn);
    ng = n_geom;
  }
  V3 world_to_local(const V3& v) const { return V3(dot(v, sn), dot(v, tn), dot(v, nn)); }
  // :132-149.  `strict_flags` reproduces the as-written matches_flags (D7 -> always black).
  RGB f(const V3& wo_w, const V3& wi_w, bool strict_flags) const {
    uint32_t flags = (dot(wi_w, ng) * dot(wo_w, ng) > 0.0f) ? (BSDF_ALL & ~BSDF_TRANSMISSION)
                                                             : (BSDF_ALL & ~BSDF_REFLECTION);
    V3 wo = world_to_local(wo_w), wi = world_to_local(wi_w);
    RGB acc(0.0f);
    for (int i = 0; i < n_bxdfs; ++i) {
      uint32_t ty = bxdfs[i].type();
      bool match = strict_flags ? ((ty & flags) == flags)   // bxdf_type.contains(flags)
                                : ((ty & flags) == ty);     // pbrt semantics
      if (match) acc = acc + bxdfs[i].f(wo, wi);
    }
    return acc;
  }
};

// ---------------------------------------------------------------------------------------------
// Lights.  kind 0 = PointLight (light/point.rs), 1 = SpotLight (light/spot.rs),
// 2 = diffuse area light over emissive triangles (ORACLE-DEFINED, SURVEY D9/A13: the reference's
// AreaLight is a stub; pbrt-v2 DiffuseAreaLight semantics; parity unpinned).
struct Light {
  int kind = 0;
  V3 pos;
  RGB intensity;  // point/spot: I ; area: emitted radiance L
  Transform world_to_light;
  float cos_total_width = 0.f, cos_falloff_start = 0.f;
  int num_samples = 1;
  // area light: emissive triangles in world space (refine-reversed vertex order) + area CDF
  struct Tri {
    V3 p1, p2, p3, nn;
    float area;
  };
  std::vector<Tri> tris;
  std::vector<float> cdf;  // size tris+1, cdf[0] = 0, cdf.back() = 1
  float total_area = 0.f;
};

struct ShadowQuery {  // visibility_tester.rs:16-24
  Ray ray;
};
inline Ray vis_segment(const V3& p1, float eps1, const V3& p2, float eps2, float time) {
  float dist = distance(p1, p2);
  V3 dir = (p2 - p1) / dist;
  Ray r(p1, dir, eps1);
  r.maxt = (1.0f - eps2) * dist;
  r.time = time;
  return r;
}

// Returns false when the sample contributes nothing before visibility (li black or pdf == 0).
inline bool light_sample_l(const Light& lt, const V3& p, float p_eps, float u1, float u2,
                           float time, RGB* li, V3* wi, float* pdf, Ray* vis) {
  if (lt.kind == 0) {  // point.rs:28-35 (D17: wi un-normalised)
    V3 w = lt.pos - p;
    *pdf = 1.0f;
    *vis = vis_segment(p, p_eps, lt.pos, 0.0f, time);
    *li = lt.intensity / length_squared(w);
    *wi = w;
  } else if (lt.kind == 1) {  // spot.rs:37-65
    V3 w = normalize(lt.pos - p);
    *pdf = 1.0f;
    *vis = vis_segment(p, p_eps, lt.pos, 0.0f, time);
    V3 wl = lt.world_to_light.vec(-w);
    float cos_theta = wl.z, fall;
    if (cos_theta < lt.cos_total_width)
      fall = 0.0f;
    else if (cos_theta > lt.cos_falloff_start)
      fall = 1.0f;
    else {
      float delta = (cos_theta - lt.cos_total_width) / (lt.cos_falloff_start - lt.cos_total_width);
      fall = delta * delta * delta * delta;
    }
    RGB i = lt.intensity * fall;
    *li = i / length_squared(w);
    *wi = w;
  } else {
    // ORACLE-DEFINED area light sampling (A13): triangle k by area CDF on u1 (u1 re-stretched),
    // uniform point b0 = 1 - sqrt(u1'), b1 = u2 * sqrt(u1'); pdf wrt solid angle.
    size_t k = 0;
    while (k + 1 < lt.tris.size() && u1 >= lt.cdf[k + 1]) ++k;
    float u1p = (u1 - lt.cdf[k]) / (lt.cdf[k + 1] - lt.cdf[k]);
    float su = std::sqrt(u1p);
    float b0 = 1.0f - su, b1 = u2 * su;
    const Light::Tri& t = lt.tris[k];
    V3 ps = b0 * t.p1 + b1 * t.p2 + (1.0f - b0 - b1) * t.p3;
    V3 w = normalize(ps - p);
    float cos_l = dot(t.nn, -w);
    float d2 = length_squared(ps - p);
    *wi = w;
    *vis = vis_segment(p, p_eps, ps, 1e-3f, time);
    if (!(cos_l > 0.0f)) {
      *li = RGB(0.0f);
      *pdf = 0.0f;
    } else {
      *li = lt.intensity;
      *pdf = d2 / (std::fabs(cos_l) * lt.total_area);
    }
  }
  return !(li->is_black() || *pdf == 0.0f);
}

// ---------------------------------------------------------------------------------------------
struct Scene {
  Geometry geom;
  BVH bvh;
  std::vector<Material> materials;
  TextureTable textures;
  std::vector<Light> lights;
  bool strict_flags = false;  // Appendix C `as_written_flags`
  uint32_t max_depth = 1;

  // Total light-sample pairs per camera sample (D11 extension).
  int light_sample_pairs() const {
    int s = 0;
    for (const Light& l : lights)
      if (l.kind == 2) s += l.num_samples;
    return s;
  }
};

// Full surface interaction at a final hit: Triangle::intersect dg part (mesh.rs:220-262),
// Sphere::intersect dg part (sphere.rs:143-180 + helpers.rs:17-43), compute_differentials,
// get_shading_geometry (mesh.rs:105-193).
struct SurfacePoint {
  DiffGeom dg;   // geometric, with differentials
  DiffGeom dgs;  // shading
  float ray_epsilon = 0.f;
  const Prim* prim = nullptr;
};

inline void tri_uvs(const Prim& pr, float uv[3][2]) {  // mesh.rs:74-87
  const Mesh& m = *pr.mesh;
  if (!m.uvs.empty()) {
    for (int k = 0; k < 3; ++k) {
      uv[k][0] = m.uvs[2 * pr.v[k]];
      uv[k][1] = m.uvs[2 * pr.v[k] + 1];
    }
  } else {
    uv[0][0] = 0.f; uv[0][1] = 0.f;
    uv[1][0] = 1.f; uv[1][1] = 0.f;
    uv[2][0] = 1.f; uv[2][1] = 1.f;
  }
}

inline DiffGeom tri_dg(const Prim& pr, const Ray& r, float t, float b1, float b2) {
  const Mesh& m = *pr.mesh;
  const V3 &p1 = m.p[pr.v[0]], &p2 = m.p[pr.v[1]], &p3 = m.p[pr.v[2]];
  float uvs[3][2];
  tri_uvs(pr, uvs);
  float du1 = uvs[0][0] - uvs[2][0];
  float du2 = uvs[1][0] - uvs[2][0];
  float dv1 = uvs[0][1] - uvs[2][1];
  float dv2 = uvs[1][1] - uvs[2][1];
  V3 dp1 = p1 - p3, dp2 = p2 - p3;
  V3 dpdu, dpdv;
  float determinant = du1 * dv2 - dv1 * du2;
  if (determinant == 0.0f) {
    coordinate_system(normalize(cross(p3 - p1, p2 - p1)), &dpdu, &dpdv);
  } else {
    float inv_det = 1.0f / determinant;
    dpdu = (dv2 * dp1 - dv1 * dp2) * inv_det;
    dpdv = (-du2 * dp1 + du1 * dp2) * inv_det;
  }
  float b0 = 1.0f - b1 - b2;
  float tu = b0 * uvs[0][0] + b1 * uvs[1][0] + b2 * uvs[2][0];
  float tv = b0 * uvs[0][1] + b1 * uvs[1][1] + b2 * uvs[2][1];
  return DiffGeom(r.at(t), dpdu, dpdv, V3(), V3(), tu, tv, &m.base);
}

// helpers.rs:17-43
inline DiffGeom compute_dg(const ShapeBase& base, float u, float v, const V3& p_hit, const V3& dpdu, const V3& dpdv,
                           const V3& d2pduu, const V3& d2pduv, const V3& d2pdvv) {
  float ee = dot(dpdu, dpdu), ff = dot(dpdu, dpdv), gg = dot(dpdv, dpdv);
  V3 nn = normalize(cross(dpdu, dpdv));
  float e = dot(nn, d2pduu), f = dot(nn, d2pduv), g = dot(nn, d2pdvv);
  float inveeggff2 = 1.0f / (ee * gg - ff * ff);
  V3 dndu = (f * ff - e * gg) * inveeggff2 * dpdu + (e * ff - f * ee) * inveeggff2 * dpdv;
  V3 dndv = (g * ff - f * gg) * inveeggff2 * dpdu + (f * ff - g * ee) * inveeggff2 * dpdv;
  const Transform& o2w = base.o2w;
  return DiffGeom(o2w.pt(p_hit), o2w.vec(dpdu), o2w.vec(dpdv), o2w.nrm(dndu), o2w.nrm(dndv), u, v, &base);
}

// Cylinder::intersect dg part (cylinder.rs:127-154)
inline DiffGeom cylinder_dg(const Sphere& s, const Ray& world_ray, float t_hit, float phi) {
  Ray ray = xf_ray(s.base.w2o, world_ray);
  V3 p_hit = ray.at(t_hit);
  float u = phi / s.phi_max;
  float v = (p_hit.z - s.z_min) / (s.z_max - s.z_min);
  V3 dpdu = s.phi_max * V3(-p_hit.y, p_hit.x, 0.0f);
  V3 dpdv(0.0f, 0.0f, s.z_max - s.z_min);
  V3 d2pduu = -s.phi_max * s.phi_max * V3(p_hit.x, p_hit.y, 0.0f);
  return compute_dg(s.base, u, v, p_hit, dpdu, dpdv, d2pduu, V3(), V3());
}

// Disk::intersect dg part (disk.rs:107-133); nn is overwritten from the OBJECT-space ray origin's
// z against 0 (not against the disk height), exactly as written.
inline DiffGeom disk_dg(const Sphere& s, const Ray& world_ray, float t_hit, float phi) {
  Ray ray = xf_ray(s.base.w2o, world_ray);
  V3 p_hit = ray.at(t_hit);
  float u = phi / s.phi_max;
  float dist = std::sqrt(p_hit.x * p_hit.x + p_hit.y * p_hit.y);
  float v = 1.0f - (dist - s.inner_radius) / (s.radius - s.inner_radius);
  V3 dpdu = (s.phi_max / (2.0f * PI_F)) * V3(-s.phi_max * p_hit.y, s.phi_max * p_hit.x, 0.0f);
  V3 dpdv = ((s.inner_radius - s.radius) / dist) * V3(p_hit.x, p_hit.y, 0.0f);
  const Transform& o2w = s.base.o2w;
  DiffGeom dg(o2w.pt(p_hit), o2w.vec(dpdu), o2w.vec(dpdv), o2w.nrm(V3()), o2w.nrm(V3()), u, v, &s.base);
  dg.nn = ray.o.z > 0.0f ? o2w.nrm(V3(0.0f, 0.0f, 1.0f)) : o2w.nrm(V3(0.0f, 0.0f, -1.0f));
  return dg;
}

inline DiffGeom sphere_dg(const Sphere& s, const Ray& world_ray, float t_hit, float phi) {
  if (s.shape == 1) return cylinder_dg(s, world_ray, t_hit, phi);
  if (s.shape == 2) return disk_dg(s, world_ray, t_hit, phi);
  Ray ray = xf_ray(s.base.w2o, world_ray);
  V3 p_hit = ray.at(t_hit);
  float u = phi / s.phi_max;
  float theta = std::acos(rclamp(p_hit.z / s.radius, -1.0f, 1.0f));
  float v = (theta - s.theta_min) / (s.theta_max - s.theta_min);
  float zradius = std::sqrt(p_hit.x * p_hit.x + p_hit.y * p_hit.y);
  float inv_zradius = 1.0f / zradius;
  float cos_phi = p_hit.x * inv_zradius;
  float sin_phi = p_hit.y * inv_zradius;
  V3 dpdu(-s.phi_max * p_hit.y, s.phi_max * p_hit.x, 0.0f);
  V3 dpdv = (s.theta_max - s.theta_min) *
            V3(p_hit.z * cos_phi, p_hit.z * sin_phi, -s.radius * std::sin(theta));
  V3 d2pduu = -s.phi_max * s.phi_max * V3(p_hit.x, p_hit.y, 0.0f);
  V3 d2pduv = (s.theta_max - s.theta_min) * p_hit.z * s.phi_max * V3(-sin_phi, cos_phi, 0.0f);
  V3 d2pdvv = -(s.theta_max - s.theta_min) * (s.theta_max - s.theta_min) * p_hit;
  // helpers.rs:17-43
  float ee = dot(dpdu, dpdu), ff = dot(dpdu, dpdv), gg = dot(dpdv, dpdv);
  V3 nn = normalize(cross(dpdu, dpdv));
  float e = dot(nn, d2pduu), f = dot(nn, d2pduv), g = dot(nn, d2pdvv);
  float inveeggff2 = 1.0f / (ee * gg - ff * ff);
  V3 dndu = (f * ff - e * gg) * inveeggff2 * dpdu + (e * ff - f * ee) * inveeggff2 * dpdv;
  V3 dndv = (g * ff - f * gg) * inveeggff2 * dpdu + (f * ff - g * ee) * inveeggff2 * dpdv;
  const Transform& o2w = s.base.o2w;
  return DiffGeom(o2w.pt(p_hit), o2w.vec(dpdu), o2w.vec(dpdv), o2w.nrm(dndu), o2w.nrm(dndv), u, v,
                  &s.base);
}

// mesh.rs:105-193
inline DiffGeom tri_shading_geometry(const Prim& pr, const DiffGeom& dg) {
  const Mesh& m = *pr.mesh;
  if (m.n.empty() && m.s.empty()) return dg;
  const Transform& o2w = m.base.o2w;
  float uv[3][2];
  tri_uvs(pr, uv);
  float b[3];
  {
    float a[2][2] = {{uv[1][0] - uv[0][0], uv[2][0] - uv[0][0]},
                     {uv[1][1] - uv[0][1], uv[2][1] - uv[0][1]}};
    float c[2] = {dg.u - uv[0][0], dg.v - uv[0][1]};
    float x0, x1;
    if (solve_linear_system_2x2(a, c, &x0, &x1)) {
      b[0] = 1.0f - x0 - x1;
      b[1] = x0;
      b[2] = x1;
    } else {
      float third = 1.f / 3.f;
      b[0] = b[1] = b[2] = third;
    }
  }
  V3 ns = !m.n.empty()
              ? normalize(o2w.vec(b[0] * m.n[pr.v[0]] + b[1] * m.n[pr.v[1]] + b[2] * m.n[pr.v[2]]))
              : dg.nn;
  V3 ss = !m.s.empty()
              ? normalize(o2w.vec(b[0] * m.s[pr.v[0]] + b[1] * m.s[pr.v[1]] + b[2] * m.s[pr.v[2]]))
              : normalize(dg.dpdu);
  V3 ts = cross(ss, ns);
  if (length_squared(ts) > 0.f) {
    ss = normalize(ts);     // as written: the tuple (ts.normalize(), ns x ts) is bound to (ss, ts)
    ts = cross(ns, ts);
  } else {
    coordinate_system(ns, &ss, &ts);
  }
  V3 dndu, dndv;
  if (!m.n.empty()) {
    float du1 = uv[0][0] - uv[2][0], du2 = uv[1][0] - uv[2][0];
    float dv1 = uv[0][1] - uv[2][1], dv2 = uv[1][1] - uv[2][1];
    V3 dn1 = m.n[pr.v[0]] - m.n[pr.v[2]];
    V3 dn2 = m.n[pr.v[1]] - m.n[pr.v[2]];
    float determinant = du1 * dv2 - dv1 * du2;
    if (determinant != 0.0f) {
      float inv_det = 1.0f / determinant;
      dndu = (dv2 * dn1 - dv1 * dn2) * inv_det;
      dndv = (-du2 * dn1 + du1 * dn2) * inv_det;
    }
  }
  return DiffGeom(dg.p, ss, ts, o2w.nrm(dndu), o2w.nrm(dndv), dg.u, dg.v, &m.base);
}

struct ShadeCounters {
  uint64_t shadow_rays = 0;
  TraceCounters shadow_trace;  // early-exit any-hit counts (roofline definition)
};

// Oracle-defined Le for emissive triangles (A13): L if n . w > 0.
inline RGB emitted(const Scene& sc, const Prim& pr, const V3& nn, const V3& w) {
  if (pr.kind != Prim::TRI || pr.mesh->area_light < 0) return RGB(0.0f);
  const Light& l = sc.lights[(size_t)pr.mesh->area_light];
  return dot(nn, w) > 0.0f ? l.intensity : RGB(0.0f);
}

// WhittedIntegrator::li (integrator/whitted.rs:30-77) for a ray that hit `hit`; light_u holds the
// D11 light-sample floats for this camera sample (2 per area-light sample, in light order).
inline RGB whitted_li(const Scene& sc, const RayDifferential& rayd, const Hit& hit,
                      const float* light_u, ShadeCounters* cnt) {
  const Prim& pr = sc.bvh.prims[hit.prim];
  const Ray& ray = rayd.ray;
  // Intersection::get_bsdf (intersection.rs:40-47)
  DiffGeom dg = pr.kind == Prim::TRI ? tri_dg(pr, ray, hit.t, hit.b1, hit.b2)
                                     : sphere_dg(*pr.sphere, ray, hit.t, hit.b1);
  float ray_epsilon = hit.t * 5e-4f;  // mesh.rs:262 / sphere.rs:180
  dg.compute_differentials(rayd);
  DiffGeom dgs = pr.kind == Prim::TRI ? tri_shading_geometry(pr, dg) : dg;
  const Material& mat = sc.materials[pr.material()];
  if (mat.bump >= 0) dgs = bump_dg(sc.textures, mat.bump, dg, dgs);  // matte.rs:33-37, plastic.rs:35-39
  BSDF bsdf(dgs, dg.nn);
  if (mat.kind == 0) {  // matte.rs:30-51
    RGB r = sc.textures.eval(mat.kd, dgs).clamp(0.0f, F32_MAX);
    float sig = rclamp(sc.textures.eval(mat.sigma, dgs).c[0], 0.0f, 90.0f);
    BxDF b;
    b.r = r;
    if (sig == 0.0f) {
      b.kind = 0;
    } else {  // orennayar.rs:16-29
      b.kind = 1;
      float sigma = as_radians(sig);
      float sigma2 = sigma * sigma;
      b.a = 1.0f - (sigma2 / (2.0f * (sigma + 0.33f)));
      b.b = 0.45f * sigma2 / (sigma2 + 0.09f);
    }
    bsdf.bxdfs[bsdf.n_bxdfs++] = b;
  } else {  // plastic.rs:32-55
    RGB kd = sc.textures.eval(mat.kd, dgs).clamp(0.0f, F32_MAX);
    RGB ks = sc.textures.eval(mat.ks, dgs).clamp(0.0f, F32_MAX);
    float rough = sc.textures.eval(mat.roughness, dgs).c[0];
    float e = 1.0f / rough;
    if (e > 1000.0f || std::isnan(e)) e = 1000.0f;  // microfacet.rs:18-24
    BxDF d;
    d.kind = 0;
    d.r = kd;
    BxDF s;
    s.kind = 2;
    s.r = ks;
    s.a = e;
    bsdf.bxdfs[bsdf.n_bxdfs++] = d;
    bsdf.bxdfs[bsdf.n_bxdfs++] = s;
  }
  const V3& p = bsdf.dg_shading.p;
  const V3& n = bsdf.dg_shading.nn;
  V3 wo = -ray.d;
  RGB L = emitted(sc, pr, dg.nn, wo);  // isect.le(wo): 0 in the reference (intersection.rs:58)
  size_t lu = 0;
  for (const Light& lt : sc.lights) {
    int ns = lt.kind == 2 ? lt.num_samples : 1;
    RGB Ld(0.0f);
    bool any = false;
    for (int s = 0; s < ns; ++s) {
      float u1 = 0.f, u2 = 0.f;
      if (lt.kind == 2) {
        u1 = light_u[lu++];
        u2 = light_u[lu++];
      }
      RGB li;
      V3 wi;
      float pdf;
      Ray vis;
      if (!light_sample_l(lt, p, ray_epsilon, u1, u2, ray.time, &li, &wi, &pdf, &vis)) continue;
      RGB f = bsdf.f(wo, wi, sc.strict_flags);
      if (f.is_black()) continue;
      if (cnt) cnt->shadow_rays++;
      if (cnt) {  // count the early-exit traversal for the roofline, answer with the same boolean
        Ray probe = vis;
        (void)sc.bvh.intersect_p(probe, true, &cnt->shadow_trace);
      }
      if (sc.bvh.intersect_p(vis, false)) continue;  // !visibility.unoccluded(scene)
      // whitted.rs:60-63 with T = 1 (D3)
      RGB c = f * li * abs_dot(wi, n) * RGB(1.0f) / pdf;
      if (lt.kind == 2) {
        Ld = Ld + c;
        any = true;
      } else {
        L = L + c;
      }
    }
    if (any) L = L + Ld / (float)ns;  // D10: average over the light's samples
  }
  // D8: matte/plastic have no specular lobe -> specular_reflect/transmit contribute 0.
  return L;
}

}  // namespace orc
