// Shading arithmetic of k_shade (sm_100a): DifferentialGeometry::{new_with, compute_differentials}
// (src/diff_geom.rs:52-152), Triangle::intersect's dg part and get_shading_geometry
// (src/shape/mesh.rs:105-193, 220-262), the dg of Sphere / Cylinder / Disk hits (src/shape/{sphere,cylinder,
// disk}.rs + helpers.rs:17-43), the BxDFs and BSDF::f (src/bsdf/*.rs) and VisibilityTester::segment
// (src/visibility_tester.rs:16-24), operation order as written.  Plain arithmetic on plain structs:
// the same source compiles as host code (PB_HOST_CHECK, tests/devsrc/) so the CPU test-suite can run
// it against the oracle; the product runs it on the GPU.
#pragma once
#include "../../include/pbrtb200.h"
#include "scene.cuh"      // DScene (per-triangle attribute arrays)
#include "shade_tex.cuh"  // DG

// diff_geom.rs:52-79
PB_DEV DG dg_new(f3 p, f3 dpdu, f3 dpdv, f3 dndu, f3 dndv, float u, float v, bool flip) {
  DG g;
  f3 norm = normalize3(cross3(dpdu, dpdv));
  if (flip) norm = norm * -1.f;
  g.p = p;
  g.nn = norm;
  g.u = u;
  g.v = v;
  g.dpdu = dpdu;
  g.dpdv = dpdv;
  g.dndu = dndu;
  g.dndv = dndv;
  g.dpdx = mk3(0, 0, 0);
  g.dpdy = mk3(0, 0, 0);
  g.dudx = g.dudy = g.dvdx = g.dvdy = 0.f;
  return g;
}

// diff_geom.rs:81-152 (has_differentials is always true for camera rays, camera/mod.rs:215)
PB_DEV void dg_compute_differentials(DG& g, f3 rxo, f3 ryo, f3 rxd, f3 ryd) {
  const f3 nvec = g.nn;
  const float d = -(dot3(nvec, g.p));
  f3 px, py;
  {
    const float ndrx = -(dot3(nvec, rxo) + d);
    const float ndrd = dot3(nvec, rxd);
    const float tx = ndrx / ndrd;
    px = rxo + tx * rxd;
  }
  {
    const float ndry = -(dot3(nvec, ryo) + d);
    const float ndrd = dot3(nvec, ryd);
    const float ty = ndry / ndrd;
    py = ryo + ty * ryd;
  }
  g.dpdx = px - g.p;
  g.dpdy = py - g.p;
  int ax0, ax1;
  if (fabsf(g.nn.x) > fabsf(g.nn.y) && fabsf(g.nn.x) > fabsf(g.nn.z)) {
    ax0 = 1;
    ax1 = 2;
  } else if (fabsf(g.nn.y) > fabsf(g.nn.z)) {
    ax0 = 0;
    ax1 = 2;
  } else {
    ax0 = 0;
    ax1 = 1;
  }
  const float a00 = comp(g.dpdu, ax0), a01 = comp(g.dpdv, ax0);
  const float a10 = comp(g.dpdu, ax1), a11 = comp(g.dpdv, ax1);
  if (!solve2x2_(a00, a01, a10, a11, comp(g.dpdx, ax0), comp(g.dpdx, ax1), &g.dudx, &g.dvdx))
    g.dudx = g.dvdx = 0.f;
  if (!solve2x2_(a00, a01, a10, a11, comp(g.dpdy, ax0), comp(g.dpdy, ax1), &g.dudy, &g.dvdy))
    g.dudy = g.dvdy = 0.f;
}

struct TriData {
  f3 p1, p2, p3;
  uint32_t mesh, attr;
  float uv[3][2];
};
// mesh.rs:220-262
PB_DEV DG tri_dg(const TriData& t, f3 o, f3 d, float th, float b1, float b2, bool flip) {
  const float du1 = t.uv[0][0] - t.uv[2][0];
  const float du2 = t.uv[1][0] - t.uv[2][0];
  const float dv1 = t.uv[0][1] - t.uv[2][1];
  const float dv2 = t.uv[1][1] - t.uv[2][1];
  const f3 dp1 = t.p1 - t.p3, dp2 = t.p2 - t.p3;
  f3 dpdu, dpdv;
  const float determinant = du1 * dv2 - dv1 * du2;
  if (determinant == 0.0f) {
    coordinate_system_(normalize3(cross3(t.p3 - t.p1, t.p2 - t.p1)), &dpdu, &dpdv);
  } else {
    const float inv_det = 1.0f / determinant;
    dpdu = (dv2 * dp1 - dv1 * dp2) * inv_det;
    dpdv = (-du2 * dp1 + du1 * dp2) * inv_det;
  }
  const float b0 = 1.0f - b1 - b2;
  const float tu = b0 * t.uv[0][0] + b1 * t.uv[1][0] + b2 * t.uv[2][0];
  const float tv = b0 * t.uv[0][1] + b1 * t.uv[1][1] + b2 * t.uv[2][1];
  return dg_new(o + (d * th), dpdu, dpdv, mk3(0, 0, 0), mk3(0, 0, 0), tu, tv, flip);
}

// mesh.rs:105-193 (as written, including the (ss, ts) tuple binding and the zeroed differentials)
PB_DEV DG tri_shading_geometry(const DScene& sc, const TriData& t, const pbrtb200_mesh& m,
                               const DG& dg) {
  const bool has_n = sc.tri_n && m.has_n, has_s = sc.tri_s && m.has_s;
  if (!has_n && !has_s) return dg;
  float b[3];
  {
    float x0, x1;
    if (solve2x2_(t.uv[1][0] - t.uv[0][0], t.uv[2][0] - t.uv[0][0], t.uv[1][1] - t.uv[0][1],
                  t.uv[2][1] - t.uv[0][1], dg.u - t.uv[0][0], dg.v - t.uv[0][1], &x0, &x1)) {
      b[0] = 1.0f - x0 - x1;
      b[1] = x0;
      b[2] = x1;
    } else {
      const float third = 1.f / 3.f;
      b[0] = b[1] = b[2] = third;
    }
  }
  f3 n0, n1, n2;
  if (has_n) {
    const float* q = sc.tri_n + 9ull * t.attr;
    n0 = mk3(q[0], q[1], q[2]);
    n1 = mk3(q[3], q[4], q[5]);
    n2 = mk3(q[6], q[7], q[8]);
  }
  f3 ns = has_n ? normalize3(xf_vec(m.o2w, b[0] * n0 + b[1] * n1 + b[2] * n2)) : dg.nn;
  f3 ss;
  if (has_s) {
    const float* q = sc.tri_s + 9ull * t.attr;
    ss = normalize3(xf_vec(
        m.o2w, b[0] * mk3(q[0], q[1], q[2]) + b[1] * mk3(q[3], q[4], q[5]) + b[2] * mk3(q[6], q[7], q[8])));
  } else {
    ss = normalize3(dg.dpdu);
  }
  f3 ts = cross3(ss, ns);
  if (len2(ts) > 0.f) {
    ss = normalize3(ts);
    ts = cross3(ns, ts);
  } else {
    coordinate_system_(ns, &ss, &ts);
  }
  f3 dndu = mk3(0, 0, 0), dndv = mk3(0, 0, 0);
  if (has_n) {
    const float du1 = t.uv[0][0] - t.uv[2][0], du2 = t.uv[1][0] - t.uv[2][0];
    const float dv1 = t.uv[0][1] - t.uv[2][1], dv2 = t.uv[1][1] - t.uv[2][1];
    const f3 dn1 = n0 - n2, dn2 = n1 - n2;
    const float determinant = du1 * dv2 - dv1 * du2;
    if (determinant != 0.0f) {
      const float inv_det = 1.0f / determinant;
      dndu = (dv2 * dn1 - dv1 * dn2) * inv_det;
      dndv = (-du2 * dn1 + du1 * dn2) * inv_det;
    }
  }
  return dg_new(dg.p, ss, ts, xf_nrm(m.o2w_inv, dndu), xf_nrm(m.o2w_inv, dndv), dg.u, dg.v,
                m.flip != 0);
}

// sphere.rs:143-180 + helpers.rs:17-43
PB_DEV DG sphere_dg(const pbrtb200_sphere80& s, const float* o2w, f3 ow, f3 dw, float t_hit,
                    float phi) {
  const f3 o = xf_pt(s.w2o, ow), d = xf_vec(s.w2o, dw);
  const f3 p_hit = o + (d * t_hit);
  const float u = phi / s.phi_max;
  const bool flip = (s.flip & 1u) != 0;
  const uint32_t kind = (s.flip >> PBRTB200_QUADRIC_KIND_SHIFT) & 3u;
  if (kind == PBRTB200_QUADRIC_DISK) {  // disk.rs:107-133 (z_min = height, theta_min = inner radius)
    const float inner = s.theta_min;
    const float dist = sqrtf(p_hit.x * p_hit.x + p_hit.y * p_hit.y);
    const float v = 1.0f - (dist - inner) / (s.radius - inner);
    const f3 dpdu = (s.phi_max / (2.0f * PB_PI)) * mk3(-s.phi_max * p_hit.y, s.phi_max * p_hit.x, 0.0f);
    const f3 dpdv = ((inner - s.radius) / dist) * mk3(p_hit.x, p_hit.y, 0.0f);
    const f3 zero = mk3(0.f, 0.f, 0.f);
    DG g = dg_new(xf_pt(o2w, p_hit), xf_vec(o2w, dpdu), xf_vec(o2w, dpdv), xf_nrm(s.w2o, zero),
                  xf_nrm(s.w2o, zero), u, v, flip);
    // nn is overwritten from the object-space ray origin's z against 0, as written (disk.rs:127-131)
    g.nn = xf_nrm(s.w2o, mk3(0.0f, 0.0f, o.z > 0.0f ? 1.0f : -1.0f));
    return g;
  }
  if (kind == PBRTB200_QUADRIC_CYLINDER) {  // cylinder.rs:127-154 + helpers.rs:17-43
    const float v = (p_hit.z - s.z_min) / (s.z_max - s.z_min);
    const f3 dpdu = s.phi_max * mk3(-p_hit.y, p_hit.x, 0.0f);
    const f3 dpdv = mk3(0.0f, 0.0f, s.z_max - s.z_min);
    const f3 d2pduu = -s.phi_max * s.phi_max * mk3(p_hit.x, p_hit.y, 0.0f);
    const f3 zero = mk3(0.f, 0.f, 0.f);
    const float ee = dot3(dpdu, dpdu), ff = dot3(dpdu, dpdv), gg = dot3(dpdv, dpdv);
    const f3 nn = normalize3(cross3(dpdu, dpdv));
    const float e = dot3(nn, d2pduu), f = dot3(nn, zero), g = dot3(nn, zero);
    const float inveeggff2 = 1.0f / (ee * gg - ff * ff);
    const f3 dndu = (f * ff - e * gg) * inveeggff2 * dpdu + (e * ff - f * ee) * inveeggff2 * dpdv;
    const f3 dndv = (g * ff - f * gg) * inveeggff2 * dpdu + (f * ff - g * ee) * inveeggff2 * dpdv;
    return dg_new(xf_pt(o2w, p_hit), xf_vec(o2w, dpdu), xf_vec(o2w, dpdv), xf_nrm(s.w2o, dndu),
                  xf_nrm(s.w2o, dndv), u, v, flip);
  }
  const float theta = acosf(rclampf(p_hit.z / s.radius, -1.0f, 1.0f));
  const float v = (theta - s.theta_min) / (s.theta_max - s.theta_min);
  const float zradius = sqrtf(p_hit.x * p_hit.x + p_hit.y * p_hit.y);
  const float inv_zradius = 1.0f / zradius;
  const float cos_phi = p_hit.x * inv_zradius;
  const float sin_phi = p_hit.y * inv_zradius;
  const f3 dpdu = mk3(-s.phi_max * p_hit.y, s.phi_max * p_hit.x, 0.0f);
  const f3 dpdv = (s.theta_max - s.theta_min) *
                  mk3(p_hit.z * cos_phi, p_hit.z * sin_phi, -s.radius * sinf(theta));
  const f3 d2pduu = -s.phi_max * s.phi_max * mk3(p_hit.x, p_hit.y, 0.0f);
  const f3 d2pduv = (s.theta_max - s.theta_min) * p_hit.z * s.phi_max * mk3(-sin_phi, cos_phi, 0.0f);
  const f3 d2pdvv = -(s.theta_max - s.theta_min) * (s.theta_max - s.theta_min) * p_hit;
  const float ee = dot3(dpdu, dpdu), ff = dot3(dpdu, dpdv), gg = dot3(dpdv, dpdv);
  const f3 nn = normalize3(cross3(dpdu, dpdv));
  const float e = dot3(nn, d2pduu), f = dot3(nn, d2pduv), g = dot3(nn, d2pdvv);
  const float inveeggff2 = 1.0f / (ee * gg - ff * ff);
  const f3 dndu = (f * ff - e * gg) * inveeggff2 * dpdu + (e * ff - f * ee) * inveeggff2 * dpdv;
  const f3 dndv = (g * ff - f * gg) * inveeggff2 * dpdu + (f * ff - g * ee) * inveeggff2 * dpdv;
  // Normal transform uses (o2w).m_inv == w2o
  return dg_new(xf_pt(o2w, p_hit), xf_vec(o2w, dpdu), xf_vec(o2w, dpdv), xf_nrm(s.w2o, dndu),
                xf_nrm(s.w2o, dndv), u, v, flip);
}

// ---- BxDFs (local shading frame) ---------------------------------------------------------------
PB_DEV float abs_cos_theta_(f3 v) { return fabsf(v.z); }
PB_DEV float sin_theta_(f3 v) { return sqrtf(fmaxf(0.f, 1.0f - v.z * v.z)); }  // bsdf/utils.rs:7-8
PB_DEV float cos_phi_(f3 v) {
  const float st = sin_theta_(v);
  return st == 0.0f ? 1.0f : rclampf(v.x / st, -1.0f, 1.0f);
}
PB_DEV float sin_phi_(f3 v) {
  const float st = sin_theta_(v);
  return st == 0.0f ? 0.0f : rclampf(v.y / st, -1.0f, 1.0f);
}
PB_DEV f3 mul3(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
PB_DEV f3 div3s(f3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }  // Spectrum / f32
PB_DEV bool is_black(f3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }

// fresnel.rs:62-91 Dielectric arm (all channels equal)
PB_DEV float fresnel_dielectric_(float cosi, float eta_i, float eta_t) {
  const float ci = rclampf(cosi, -1.0f, 1.0f);
  float ei = eta_i, et = eta_t;
  if (cosi <= 0.0f) {
    const float tmp = ei;
    ei = et;
    et = tmp;
  }
  const float sint = (ei / et) * sqrtf(fmaxf(1.0f - ci * ci, 0.0f));
  if (sint >= 1.0f) return 1.0f;
  const float cost = sqrtf(fmaxf(1.0f - sint * sint, 0.0f));
  const float aci = fabsf(ci);
  const float rparl = ((et * aci) - (ei * cost)) / ((et * aci) + (ei * cost));
  const float rperp = ((ei * aci) - (et * cost)) / ((ei * aci) + (et * cost));
  return (rparl * rparl + rperp * rperp) / 2.0f;
}
PB_DEV f3 lambertian_f(f3 r) {  // lambertian.rs:20-23
  const float invpi = 1.0f / PB_PI;
  return r * invpi;
}
PB_DEV f3 orennayar_f(f3 r, float A, float B, f3 wo, f3 wi) {  // orennayar.rs:36-60
  const float sinthetai = sin_theta_(wi), sinthetao = sin_theta_(wo);
  float maxcos = 0.0f;
  if (!(sinthetai < 1e-4f || sinthetao < 1e-4f)) {
    const float sinphii = sin_phi_(wi), cosphii = cos_phi_(wi);
    const float sinphio = sin_phi_(wo), cosphio = cos_phi_(wo);
    maxcos = fmaxf(cosphii * cosphio + sinphii * sinphio, 0.0f);
  }
  float sinalpha, tanbeta;
  if (abs_cos_theta_(wi) > abs_cos_theta_(wo)) {
    sinalpha = sinthetao;
    tanbeta = sinthetai / abs_cos_theta_(wi);
  } else {
    sinalpha = sinthetai;
    tanbeta = sinthetao / abs_cos_theta_(wo);
  }
  const float invpi = 1.0f / PB_PI;
  return r * invpi * (A + B * maxcos * sinalpha * tanbeta);
}
PB_DEV f3 microfacet_blinn_f(f3 r, float e, f3 wo, f3 wi) {  // microfacet.rs:32-38,71-99
  const float cos_o = abs_cos_theta_(wo), cos_i = abs_cos_theta_(wi);
  if (cos_o == 0.0f || cos_i == 0.0f) return mk3(0.f, 0.f, 0.f);
  const f3 wh = normalize3(wo + wi);
  const float cos_h = dot3(wi, wh);
  const float F = fresnel_dielectric_(cos_h, 1.5f, 1.0f);
  const float invtwopi = 1.0f / (2.0f * PB_PI);
  const float D = (e + 2.0f) * invtwopi * powf(abs_cos_theta_(wh), e);
  const float ndotwh = abs_cos_theta_(wh), ndotwo = abs_cos_theta_(wo), ndotwi = abs_cos_theta_(wi);
  const float wodotwh = fabsf(dot3(wo, wh));
  const float G =
      fminf(fminf(2.0f * ndotwh * ndotwo / wodotwh, 2.0f * ndotwh * ndotwi / wodotwh), 1.0f);
  return div3s(mul3(r * D * G, mk3(F, F, F)), 4.0f * cos_i * cos_o);
}

struct DBSDF {
  f3 nn, ng, sn, tn;
  int kind;      // 0 matte/Lambertian, 1 matte/OrenNayar, 2 plastic (Lambertian + Blinn microfacet)
  f3 kd, ks;
  float a, b;    // OrenNayar A,B ; plastic: a = Blinn exponent
};
PB_DEV f3 bsdf_f(const DBSDF& bs, f3 wo_w, f3 wi_w, bool strict_flags) {  // bsdf/mod.rs:132-149
  const bool reflect = dot3(wi_w, bs.ng) * dot3(wo_w, bs.ng) > 0.0f;
  // D7: with the as-written matches_flags no BxDF ever matches the wide mask -> black.
  if (strict_flags || !reflect) return mk3(0.f, 0.f, 0.f);
  const f3 wo = mk3(dot3(wo_w, bs.sn), dot3(wo_w, bs.tn), dot3(wo_w, bs.nn));
  const f3 wi = mk3(dot3(wi_w, bs.sn), dot3(wi_w, bs.tn), dot3(wi_w, bs.nn));
  f3 f = mk3(0.f, 0.f, 0.f);
  if (bs.kind == 0) {
    return f + lambertian_f(bs.kd);
  } else if (bs.kind == 1) {
    f = f + orennayar_f(bs.kd, bs.a, bs.b, wo, wi);
  } else {
    f = f + lambertian_f(bs.kd);
    f = f + microfacet_blinn_f(bs.ks, bs.a, wo, wi);
  }
  return f;
}

// visibility_tester.rs:16-24
PB_DEV void vis_segment(f3 p1, float eps1, f3 p2, float eps2, pbrtb200_ray32* r) {
  const float dist = len3(p1 - p2);
  const f3 dir = (p2 - p1) / dist;
  r->o[0] = p1.x;
  r->o[1] = p1.y;
  r->o[2] = p1.z;
  r->mint = eps1;
  r->d[0] = dir.x;
  r->d[1] = dir.y;
  r->d[2] = dir.z;
  r->maxt = (1.0f - eps2) * dist;
}

