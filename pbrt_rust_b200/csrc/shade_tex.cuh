// Texture mappings, procedural textures and bump mapping of the shading stage (sm_100a).
//
// Replaces TextureMapping2D::map for UV / Spherical / Cylindrical / Planar mappings
// (src/texture/mapping2d.rs:49-210), IdentityMapping3D::map (src/texture/mapping3d.rs:57-63),
// noise / fbm / turbulence (src/texture/noise.rs:57-145), Scale / Mix / Bilerp / Dots / FBm /
// Wrinkled textures (src/texture/{mod,mix,bilerp,dots,fbm}.rs) and material::bump
// (src/material/mod.rs:23-77).  Operation order follows the reference line by line.
//
// Everything in this header is arithmetic on plain structs, so the same source also compiles as
// host code (PB_HOST_CHECK, tests/devsrc/): the CPU test-suite runs THIS source against the oracle
// where no GPU exists.  That build is test infrastructure; the product only ever runs it on the GPU.
#pragma once
#include "../../include/pbrtb200.h"
#include "dmath.cuh"

#ifdef PB_HOST_CHECK
#define PB_NOINLINE inline
#define PB_TABLE static const
#else
#define PB_NOINLINE __device__ __noinline__
#define PB_TABLE static __device__ const
#endif

struct DG {
  f3 p, nn;
  float u, v;
  f3 dpdu, dpdv, dndu, dndv, dpdx, dpdy;
  float dudx, dudy, dvdx, dvdy;
};

// ---- 2D mappings --------------------------------------------------------------------------------
PB_DEV void tex_map(const pbrtb200_texture& tx, const DG& dg, float m[6]) {
  if (tx.map_kind == PBRTB200_MAP_UV) {  // mapping2d.rs:66-76
    m[0] = tx.map[0] * dg.u + tx.map[2];
    m[1] = tx.map[1] * dg.v + tx.map[3];
    m[2] = tx.map[0] * dg.dudx;
    m[3] = tx.map[1] * dg.dvdx;
    m[4] = tx.map[0] * dg.dudy;
    m[5] = tx.map[1] * dg.dvdy;
  } else {  // mapping2d.rs:199-210
    const f3 vs = mk3(tx.map[0], tx.map[1], tx.map[2]), vt = mk3(tx.map[3], tx.map[4], tx.map[5]);
    m[0] = tx.map[6] + dot3(dg.p, vs);
    m[1] = tx.map[7] + dot3(dg.p, vt);
    m[2] = dot3(vs, dg.dpdx);
    m[3] = dot3(vt, dg.dpdx);
    m[4] = dot3(vs, dg.dpdy);
    m[5] = dot3(vt, dg.dpdy);
  }
}
PB_DEV float bump_int_(float x) {  // checkerboard.rs:62-66
  const float half_x = x / 2.0f;
  return floorf(half_x) + 2.0f * fmaxf(half_x - floorf(half_x) - 0.5f, 0.0f);
}

// SphericalMapping2D::sphere (mapping2d.rs:120-133) / CylindricalMapping2D::cylinder (:155-166)
PB_DEV void tex_circular_(const pbrtb200_texture& tx, f3 p, float* s, float* t) {
  f3 vec = xf_pt(tx.map, p);
  if (vec.x == 0.0f && vec.y == 0.0f && vec.z == 0.0f)
    vec = mk3(1.0f, 0.0f, 0.0f);
  else
    vec = normalize3(vec);
  if (tx.map_kind == PBRTB200_MAP_SPHERICAL) {  // spherical_theta / spherical_phi (vector.rs:210-217)
    const float frac_1_pi = 0.318309886183790671537767526745028724f;
    const float theta = acosf(rclampf(vec.z, -1.0f, 1.0f));
    float phi = atan2f(vec.y, vec.x);
    if (phi < 0.0f) phi = phi + 2.0f * PB_PI;
    *s = theta * frac_1_pi;
    *t = phi * frac_1_pi * 0.5f;
  } else {
    *s = (PB_PI + atan2f(vec.y, vec.x)) / (2.0f * PB_PI);
    *t = vec.z;
  }
}
PB_DEV float tex_singularity_(float res) {  // mapping2d.rs:83-91
  if (res > 0.5f) return 1.0f - res;
  if (res < -0.5f) return -(res + 1.0f);
  return res;
}
// All four 2D mappings.  get_circular_differentials: mapping2d.rs:78-104.
PB_DEV void tex_map_ext(const pbrtb200_texture& tx, const DG& dg, float m[6]) {
  if (tx.map_kind == PBRTB200_MAP_SPHERICAL || tx.map_kind == PBRTB200_MAP_CYLINDRICAL) {
    float s, t, sx, tx_, sy, ty;
    tex_circular_(tx, dg.p, &s, &t);
    const float delta = 0.1f;
    tex_circular_(tx, dg.p + delta * dg.dpdx, &sx, &tx_);
    tex_circular_(tx, dg.p + delta * dg.dpdy, &sy, &ty);
    m[0] = s;
    m[1] = t;
    m[2] = (sx - s) / delta;
    m[3] = tex_singularity_((tx_ - t) / delta);
    m[4] = (sy - s) / delta;
    m[5] = tex_singularity_((ty - t) / delta);
  } else {
    tex_map(tx, dg, m);
  }
}

// ---- noise (texture/noise.rs) -------------------------------------------------------------------
// NOISE_PERM (noise.rs:6-55) stores these 256 entries twice; an index < 512 is taken mod 256.
// Lanes index it divergently, so it lives in global memory (L1-cached) and not in the constant bank.
PB_TABLE unsigned char pb_noise_perm[256] = {
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
PB_DEV uint32_t noise_perm_(uint32_t i) { return pb_noise_perm[i & 255u]; }
PB_DEV float noise_grad_(uint32_t x, uint32_t y, uint32_t z, float dx, float dy, float dz) {  // :57-62
  const uint32_t h = noise_perm_(noise_perm_(noise_perm_(x) + y) + z) & 15u;
  const float u = (h < 8u || h == 12u || h == 13u) ? dx : dy;
  const float v = (h < 4u || h == 12u || h == 13u) ? dy : dz;
  return ((h & 1u) == 0u ? u : -u) + ((h & 2u) == 0u ? v : -v);
}
PB_DEV float noise_weight_(float t) {  // :64-68
  const float t3 = t * t * t;
  const float t4 = t3 * t;
  return 6.0f * t4 * t - 15.0f * t4 + 10.0f * t3;
}
PB_NOINLINE float noise_(float x, float y, float z) {  // noise.rs:70-103
  const uint32_t ix = (uint32_t)(f2i_sat(floorf(x)) & 255);
  const uint32_t iy = (uint32_t)(f2i_sat(floorf(y)) & 255);
  const uint32_t iz = (uint32_t)(f2i_sat(floorf(z)) & 255);
  const float dx = x - floorf(x), dy = y - floorf(y), dz = z - floorf(z);
  const float w000 = noise_grad_(ix, iy, iz, dx, dy, dz);
  const float w100 = noise_grad_(ix + 1, iy, iz, dx - 1.0f, dy, dz);
  const float w010 = noise_grad_(ix, iy + 1, iz, dx, dy - 1.0f, dz);
  const float w110 = noise_grad_(ix + 1, iy + 1, iz, dx - 1.0f, dy - 1.0f, dz);
  const float w001 = noise_grad_(ix, iy, iz + 1, dx, dy, dz - 1.0f);
  const float w101 = noise_grad_(ix + 1, iy, iz + 1, dx - 1.0f, dy, dz - 1.0f);
  const float w011 = noise_grad_(ix, iy + 1, iz + 1, dx, dy - 1.0f, dz - 1.0f);
  const float w111 = noise_grad_(ix + 1, iy + 1, iz + 1, dx - 1.0f, dy - 1.0f, dz - 1.0f);
  const float wx = noise_weight_(dx), wy = noise_weight_(dy), wz = noise_weight_(dz);
  const float x00 = lerpf_(w000, w100, wx);
  const float x10 = lerpf_(w010, w110, wx);
  const float x01 = lerpf_(w001, w101, wx);
  const float x11 = lerpf_(w011, w111, wx);
  const float y0 = lerpf_(x00, x10, wy);
  const float y1 = lerpf_(x01, x11, wy);
  return lerpf_(y0, y1, wz);
}
PB_DEV float smoothstep_(float mn, float mx, float value) {  // noise.rs:107-110
  const float v = rclampf((value - mn) / (mx - mn), 0.0f, 1.0f);
  return v * v * (-2.0f * v + 3.0f);
}
// fbm (noise.rs:112-127) / turbulence (:129-145; as written |noise| only at the partial octave)
PB_DEV float fbm_(bool turb, f3 p, f3 dpdx, f3 dpdy, float omega, int max_octaves) {
  const float s2 = fmaxf(len2(dpdx), len2(dpdy));
  const float foctaves = rclampf(-1.0f - 0.5f * log2f(s2), 0.0f, (float)max_octaves);
  const int octaves = f2i_sat(floorf(foctaves));
  float sum = 0.0f, lambda = 1.0f, o = 1.0f;
  for (int i = 0; i < octaves; ++i) {
    const float v = noise_(lambda * p.x, lambda * p.y, lambda * p.z);
    sum = sum + o * v;
    lambda = lambda * 1.99f;
    o = o * omega;
  }
  const float partial_octave = foctaves - floorf(foctaves);
  float n = noise_(lambda * p.x, lambda * p.y, lambda * p.z);
  if (turb) n = fabsf(n);
  return sum + o * smoothstep_(0.3f, 0.7f, partial_octave) * n;
}

// ---- general texture evaluation -----------------------------------------------------------------
struct TexEnv {
  const pbrtb200_texture* textures;
  const pbrtb200_mipmap* mipmaps;  // image textures: headers and texel pool
  const float4* texels;
};
PB_NOINLINE f3 mip_lookup(const float4* __restrict__ texels, const pbrtb200_mipmap* __restrict__ mmp,
                          float s, float t, float dsdx, float dtdx, float dsdy, float dtdy);  // shade_mip.cuh

// Every texture kind; parents (checkerboard, scale, mix, dots) nest PBRTB200_TEX_MAX_DEPTH deep
// (validated at upload).  One out-of-line function per level keeps the code size linear in the
// depth although a level has nine child call sites.
template <int DEPTH>
PB_NOINLINE f3 tex_eval_ext(const TexEnv& env, int id, const DG& dg) {
  const pbrtb200_texture& tx = env.textures[id];
  const int kind = tx.kind;
  if (kind == PBRTB200_TEX_CONSTANT) return mk3(tx.value[0], tx.value[1], tx.value[2]);
  if (kind == PBRTB200_TEX_FBM || kind == PBRTB200_TEX_WRINKLED) {  // fbm.rs:21-26, 42-47
    const f3 dpdx = xf_vec(tx.map, dg.dpdx), dpdy = xf_vec(tx.map, dg.dpdy);  // mapping3d.rs:57-63
    const f3 p = xf_pt(tx.map, dg.p);
    const float v = fbm_(kind == PBRTB200_TEX_WRINKLED, p, dpdx, dpdy, tx.value[0], tx.aa);
    return mk3(v, v, v);
  }
  if constexpr (DEPTH > 0) {
    if (kind == PBRTB200_TEX_SCALE) {  // texture/mod.rs:81-85
      const f3 a = tex_eval_ext<DEPTH - 1>(env, tx.tex1, dg), b = tex_eval_ext<DEPTH - 1>(env, tx.tex2, dg);
      return mk3(a.x * b.x, a.y * b.y, a.z * b.z);
    }
    if (kind == PBRTB200_TEX_MIX) {  // mix.rs:22-27
      const f3 a = tex_eval_ext<DEPTH - 1>(env, tx.tex1, dg), b = tex_eval_ext<DEPTH - 1>(env, tx.tex2, dg);
      const float amt = tex_eval_ext<DEPTH - 1>(env, tx.tex3, dg).x;
      return a * (1.0f - amt) + b * amt;
    }
  }
  float m[6];
  tex_map_ext(tx, dg, m);
  if (kind == PBRTB200_TEX_UV)  // uv.rs:20-26
    return mk3(m[0] - floorf(m[0]), m[1] - floorf(m[1]), 0.0f);
  if (kind == PBRTB200_TEX_BILERP) {  // bilerp.rs:28-35
    const f3 v00 = mk3(tx.value[0], tx.value[1], tx.value[2]), v01 = mk3(tx.value[3], tx.value[4], tx.value[5]);
    const f3 v10 = mk3(tx.value[6], tx.value[7], tx.value[8]), v11 = mk3(tx.value[9], tx.value[10], tx.value[11]);
    const f3 tmp1 = v00 * (1.0f - m[0]) + v10 * m[0];
    const f3 tmp2 = v01 * (1.0f - m[0]) + v11 * m[0];
    return tmp1 * (1.0f - m[1]) + tmp2 * m[1];
  }
  if (kind == PBRTB200_TEX_IMAGE)  // imagemap.rs:200-206
    return mip_lookup(env.texels, env.mipmaps + tx.tex1, m[0], m[1], m[2], m[3], m[4], m[5]);
  if constexpr (DEPTH > 0) {
    const float s = m[0], t = m[1];
    if (kind == PBRTB200_TEX_DOTS) {  // dots.rs:23-46
      const float s_cell = floorf(s + 0.5f), t_cell = floorf(t + 0.5f);
      bool inside = false;
      if (noise_(s_cell + 0.5f, t_cell + 0.5f, 0.5f) > 0.0f) {
        const float radius = 0.35f;
        const float max_shift = 0.5f - radius;
        const float s_center = s_cell + max_shift * noise_(s_cell + 1.5f, t_cell + 2.8f, 0.5f);
        const float t_center = t_cell + max_shift * noise_(s_cell + 4.5f, t_cell + 9.8f, 0.5f);
        const float ds = s - s_center, dt = t - t_center;
        inside = ds * ds + dt * dt < radius * radius;
      }
      return tex_eval_ext<DEPTH - 1>(env, inside ? tx.tex1 : tx.tex2, dg);
    }
    if (kind == PBRTB200_TEX_CHECKER2D) {
      // checkerboard.rs:38-44 (i32 arithmetic wraps in release builds)
      const int sum = (int)((uint32_t)f2i_sat(floorf(s)) + (uint32_t)f2i_sat(floorf(t)));
      const bool first = (sum % 2) == 0;
      bool point = tx.aa == 0;
      float ds = 0.f, dt = 0.f, s0 = 0.f, t0 = 0.f, s1 = 0.f, t1 = 0.f;
      if (!point) {
        ds = fmaxf(fabsf(m[2]), fabsf(m[4]));
        dt = fmaxf(fabsf(m[3]), fabsf(m[5]));
        s0 = s - ds;
        t0 = t - dt;
        s1 = s + ds;
        t1 = t + dt;
        if (floorf(s0) == floorf(s1) && floorf(t0) == floorf(t1)) point = true;
      }
      if (point) return tex_eval_ext<DEPTH - 1>(env, first ? tx.tex1 : tx.tex2, dg);
      const float sint = ds > 0.0f ? (bump_int_(s1) - bump_int_(s0)) / (2.0f * ds) : 0.0f;
      const float tint = dt > 0.0f ? (bump_int_(t1) - bump_int_(t0)) / (2.0f * dt) : 0.0f;
      const float area_sq = (ds > 1.0f || dt > 1.0f) ? 0.5f : sint + tint - 2.0f * sint * tint;
      const f3 a = tex_eval_ext<DEPTH - 1>(env, tx.tex1, dg), b = tex_eval_ext<DEPTH - 1>(env, tx.tex2, dg);
      return a * (1.0f - area_sq) + b * area_sq;  // Lerp::lerp_with, utils/mod.rs:20
    }
  }
  return mk3(0.f, 0.f, 0.f);
}

// ---- material::bump (material/mod.rs:23-77) -----------------------------------------------------
// `flip` = shape.reverse_orientation ^ shape.transform_swaps_handedness of the hit shape; `ng` = the
// geometric normal (dg_geom.nn).  Returns the bump-mapped shading geometry.
PB_DEV DG bump_dg_(const TexEnv& env, int d, const DG& dgs, f3 ng, bool flip) {
  DG ev = dgs;
  float du = 0.5f * (fabsf(dgs.dudx) + fabsf(dgs.dudy));
  if (du == 0.0f) du = 0.1f;
  ev.p = dgs.p + du * dgs.dpdu;
  ev.u = dgs.u + du;
  ev.nn = normalize3(cross3(dgs.dpdu, dgs.dpdv) + du * dgs.dndu);
  const float u_displace = tex_eval_ext<PBRTB200_TEX_MAX_DEPTH>(env, d, ev).x;
  float dv = 0.5f * (fabsf(dgs.dvdx) + fabsf(dgs.dvdy));
  if (dv == 0.0f) dv = 0.1f;
  ev.p = dgs.p + dv * dgs.dpdv;
  ev.u = dgs.u;
  ev.v = dgs.v + dv;
  ev.nn = normalize3(cross3(dgs.dpdu, dgs.dpdv) + dv * dgs.dndv);
  const float v_displace = tex_eval_ext<PBRTB200_TEX_MAX_DEPTH>(env, d, ev).x;
  // d.evaluate(dg_shading): through `ev` again, so that only this one copy has its address taken
  // (the caller's dgs stays in registers)
  ev.p = dgs.p;
  ev.v = dgs.v;
  ev.nn = dgs.nn;
  const float displace = tex_eval_ext<PBRTB200_TEX_MAX_DEPTH>(env, d, ev).x;
  DG b = dgs;
  b.dpdu = dgs.dpdu + (u_displace - displace) / du * dgs.nn + displace * dgs.dndu;
  b.dpdv = dgs.dpdv + (v_displace - displace) / dv * dgs.nn + displace * dgs.dndv;
  b.nn = normalize3(cross3(b.dpdu, b.dpdv));
  if (flip) b.nn = mk3(-b.nn.x, -b.nn.y, -b.nn.z);
  if (dot3(b.nn, ng) < 0.0f) b.nn = -b.nn;  // face_forward (normal.rs:22-24)
  return b;
}
