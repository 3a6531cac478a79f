// Ray / box and ray / triangle arithmetic of the traversal kernels (sm_100a): BBox::intersect
// (src/bbox.rs:185-209), Triangle::get_intersection_point (src/shape/mesh.rs:41-72) and the
// perspective camera ray (src/camera/mod.rs:168-271), operation order as written.  Plain arithmetic: the same source compiles as host code (PB_HOST_CHECK,
// tests/devsrc/) so the CPU test-suite can run it against the oracle; the product runs it on the GPU.
#pragma once
#include "scene.cuh"  // DCamera (includes dmath.cuh)

// bbox.rs:185-209.  Returns pass/fail and T0 (the entry distance after all three axes).
PB_DEV bool slab_test(float bminx, float bminy, float bminz, float bmaxx, float bmaxy, float bmaxz,
                      f3 o, f3 inv, float mint, float maxt, float* T0) {
  float t0 = mint, t1 = maxt;
  {
    float ta = (bminx - o.x) * inv.x, tb = (bmaxx - o.x) * inv.x;
    bool sw = ta > tb;  // NaN compares false: no swap, as in the reference
    float tn = sw ? tb : ta, tf = sw ? ta : tb;
    t0 = fmaxf(tn, t0);
    t1 = fminf(tf, t1);
  }
  {
    float ta = (bminy - o.y) * inv.y, tb = (bmaxy - o.y) * inv.y;
    bool sw = ta > tb;
    float tn = sw ? tb : ta, tf = sw ? ta : tb;
    t0 = fmaxf(tn, t0);
    t1 = fminf(tf, t1);
  }
  {
    float ta = (bminz - o.z) * inv.z, tb = (bmaxz - o.z) * inv.z;
    bool sw = ta > tb;
    float tn = sw ? tb : ta, tf = sw ? ta : tb;
    t0 = fmaxf(tn, t0);
    t1 = fminf(tf, t1);
  }
  *T0 = t0;
  return !(t0 > t1);
}

// mesh.rs:41-72
PB_DEV bool tri_hit(f3 p1, f3 p2, f3 p3, f3 o, f3 d, float mint, float maxt, float* t_out,
                    float* b1_out, float* b2_out) {
  f3 e1 = p2 - p1;
  f3 e2 = p3 - p1;
  f3 s1 = cross3(d, e2);
  float divisor = dot3(s1, e1);
  if (divisor == 0.f) return false;
  float inv_divisor = 1.0f / divisor;
  f3 s = o - p1;
  float b1 = dot3(s1, s) * inv_divisor;
  if (b1 < 0.0f || b1 > 1.0f) return false;
  f3 s2 = cross3(s, e1);
  float b2 = dot3(d, s2) * inv_divisor;
  if (b2 < 0.0f || (b1 + b2) > 1.0f) return false;
  float t = dot3(e2, s2) * inv_divisor;
  if (t < mint || t > maxt) return false;
  *t_out = t;
  *b1_out = b1;
  *b2_out = b2;
  return true;
}

// Slab test for rays whose 1/d is finite on every axis.  No product (b - o) * inv can then be NaN,
// so the reference's "swap if ta > tb" (bbox.rs:194-196) is exactly (min(ta,tb), max(ta,tb)):
// two FMNMX per axis instead of FSETP + 2 FSEL, and the ALU pipe is this kernel's busiest unit.
PB_DEV bool slab_test_finite(float bminx, float bminy, float bminz, float bmaxx, float bmaxy,
                             float bmaxz, f3 o, f3 inv, float mint, float maxt, float* T0) {
  const float tax = (bminx - o.x) * inv.x, tbx = (bmaxx - o.x) * inv.x;
  const float tay = (bminy - o.y) * inv.y, tby = (bmaxy - o.y) * inv.y;
  const float taz = (bminz - o.z) * inv.z, tbz = (bmaxz - o.z) * inv.z;
  const float t0 = fmaxf(fmaxf(fmaxf(fminf(tax, tbx), mint), fminf(tay, tby)), fminf(taz, tbz));
  const float t1 = fminf(fminf(fminf(fmaxf(tax, tbx), maxt), fmaxf(tay, tby)), fmaxf(taz, tbz));
  *T0 = t0;
  return !(t0 > t1);
}

// camera/mod.rs:168-195, 212-271 (Perspective arm), projective.rs:79-97, animated.rs:275-284
PB_DEV void camera_ray(const DCamera& cam, float ix, float iy, float lu, float lv, f3* o, f3* d,
                       f3* p_camera_out) {
  f3 p_camera = xf_pt44(cam.r2c, mk3(ix, iy, 0.0f));
  f3 ro = mk3(0.f, 0.f, 0.f);
  f3 rd = normalize3(p_camera);
  if (cam.lens_radius > 0.0f) {  // handle_dof (concentric_sample_disk is the identity)
    float u = lu * cam.lens_radius, v = lv * cam.lens_radius;
    float ft = cam.focal_distance / rd.z;
    f3 p_focus = ro + (rd * ft);
    ro = mk3(u, v, 0.0f);
    rd = normalize3(p_focus - ro);
  }
  *o = xf_pt44(cam.c2w, ro);
  *d = xf_vec(cam.c2w, rd);
  if (p_camera_out) *p_camera_out = p_camera;
}
