// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement (C++17) of the pbrt_rust math/geometry layer, kept line-faithful to the
// reference's floating-point operation order so that results are bit-comparable.  Build with
// `-O2 -ffp-contract=off -fno-fast-math` (see oracle/Makefile): Rust never contracts a*b+c into
// an FMA and neither may we.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this code.  Each function cites the reference file:line it follows (paths relative to the
// reference crate root).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <utility>
#include <vector>

namespace orc {

static constexpr float F32_MAX = std::numeric_limits<float>::max();
static constexpr float PI_F = 3.14159265358979323846f;  // std::f32::consts::PI

// Rust `f32::max` / `f32::min`: if one argument is NaN the other is returned (== fmaxf/fminf).
inline float rmax(float a, float b) { return std::fmax(a, b); }
inline float rmin(float a, float b) { return std::fmin(a, b); }
// Rust `f32::clamp(lo, hi)`: NaN stays NaN.
inline float rclamp(float x, float lo, float hi) {
  if (x < lo) return lo;
  if (x > hi) return hi;
  return x;
}
// Rust `f32 as i32` / `as usize`: saturating, NaN -> 0.
inline int32_t f2i(float x) {
  if (std::isnan(x)) return 0;
  if (x >= 2147483648.0f) return INT32_MAX;
  if (x <= -2147483648.0f) return INT32_MIN;
  return (int32_t)x;
}
inline uint64_t f2usize(float x) {
  if (std::isnan(x) || x <= 0.0f) return 0;
  if (x >= 18446744073709551616.0f) return UINT64_MAX;
  return (uint64_t)x;
}
// utils/mod.rs:16-21  Lerp: self*(1-t) + b*t
inline float lerpf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
// utils/mod.rs:36-39
inline float as_radians(float deg) { return deg * PI_F / 180.0f; }

// geometry/vector.rs:27-193.  Point and Normal (geometry/point.rs, geometry/normal.rs) share the
// same three-float layout and the same component-wise operator definitions, so one struct serves
// all three; the semantic differences (how a Transform applies) live in Transform below.
struct V3 {
  float x = 0.f, y = 0.f, z = 0.f;
  V3() = default;
  V3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
  float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
  float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(const V3& a, const V3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(const V3& a, const V3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(const V3& a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(const V3& a, float f) { return {a.x * f, a.y * f, a.z * f}; }
inline V3 operator*(float f, const V3& a) { return a * f; }  // vector.rs:102-110: f*v == v*f
// vector.rs:112-126: v / f == (1/f) * v
inline V3 operator/(const V3& a, float f) {
  float recip = 1.0f / f;
  return recip * a;
}
inline bool operator==(const V3& a, const V3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
// vector.rs:168-172
inline float dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float abs_dot(const V3& a, const V3& b) { return std::fabs(dot(a, b)); }
// vector.rs:174-181
inline V3 cross(const V3& a, const V3& v) {
  return {(a.y * v.z) - (a.z * v.y), (a.z * v.x) - (a.x * v.z), (a.x * v.y) - (a.y * v.x)};
}
// vector.rs:37-43
inline float length_squared(const V3& a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline float length(const V3& a) { return std::sqrt(length_squared(a)); }
// normal.rs:180-185
inline V3 normalize(const V3& a) {
  float l = length(a);
  return a / l;
}
// point.rs:17
inline float distance(const V3& a, const V3& b) { return length(a - b); }

// vector.rs:184-195 (as written: the first branch uses v1.x for BOTH components where pbrt uses
// -v1.z and v1.x; returns (v3 x v1, v3))
// normal.rs:22-24 (a NaN dot product is not < 0: the normal is kept)
inline V3 face_forward(const V3& n, const V3& v) { return dot(n, v) < 0.0f ? -n : n; }
inline void coordinate_system(const V3& v1, V3* o1, V3* o2) {
  V3 v2;
  if (std::fabs(v1.x) > std::fabs(v1.y)) {
    float inv_len = 1.0f / std::sqrt(v1.x * v1.x + v1.z * v1.z);
    v2 = V3(-v1.x * inv_len, 0.f, v1.x * inv_len);
  } else {
    float inv_len = 1.0f / std::sqrt(v1.y * v1.y + v1.z * v1.z);
    v2 = V3(0.f, v1.z * inv_len, -v1.y * inv_len);
  }
  V3 v3 = cross(v1, v2);
  *o1 = cross(v3, v1);
  *o2 = v3;
}

// ---------------------------------------------------------------------------------------------
// transform/matrix4x4.rs
struct M44 {
  float m[4][4];
  M44() {
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) m[i][j] = (i == j) ? 1.f : 0.f;
  }
  static M44 rows(float a00, float a01, float a02, float a03, float a10, float a11, float a12,
                  float a13, float a20, float a21, float a22, float a23, float a30, float a31,
                  float a32, float a33) {
    M44 r;
    float v[16] = {a00, a01, a02, a03, a10, a11, a12, a13, a20, a21, a22, a23, a30, a31, a32, a33};
    std::memcpy(r.m, v, sizeof v);
    return r;
  }
  M44 transpose() const {
    M44 r;
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) r.m[i][j] = m[j][i];
    return r;
  }
};
// matrix4x4.rs:167-181
inline M44 operator*(const M44& a, const M44& b) {
  M44 r;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j] +
                  a.m[i][3] * b.m[3][j];
  return r;
}

struct SingularMatrix : std::runtime_error {
  SingularMatrix() : std::runtime_error("Singular matrix!") {}
};

// matrix4x4.rs:66-165: LU with scaled partial pivoting, then four solves, then transpose.
inline M44 invert(const M44& a) {
  float s[4];
  for (int i = 0; i < 4; ++i) {
    float acc = 0.f;
    for (int j = 0; j < 4; ++j) acc = rmax(std::fabs(a.m[i][j]), acc);
    s[i] = acc;
  }
  M44 lu = a;
  int pivot[4] = {0, 1, 2, 3};
  for (int k = 0; k < 3; ++k) {
    float c = 0.0f;
    int p = k;
    for (int i = k; i < 4; ++i) {
      float cc = std::fabs(lu.m[i][k] / s[i]);
      if (cc > c) {
        c = cc;
        p = i;
      }
    }
    pivot[k] = p;
    if (c == 0.0f) throw SingularMatrix();
    if (p != k)
      for (int j = k; j < 4; ++j) std::swap(lu.m[k][j], lu.m[p][j]);
    for (int i = k + 1; i < 4; ++i) {
      float mi = lu.m[i][k] / lu.m[k][k];
      lu.m[i][k] = mi;
      for (int j = k + 1; j < 4; ++j) lu.m[i][j] = lu.m[i][j] - mi * lu.m[k][j];
    }
  }
  if (std::fabs(lu.m[3][3]) < 1.0e-6f) throw SingularMatrix();

  M44 cols;  // row r of `cols` = solution for unit vector e_r; result is its transpose
  for (int r = 0; r < 4; ++r) {
    float b[4] = {0.f, 0.f, 0.f, 0.f};
    b[r] = 1.0f;
    for (int k = 0; k < 3; ++k) {
      if (pivot[k] != k) std::swap(b[pivot[k]], b[k]);
      for (int i = k + 1; i < 4; ++i) b[i] = b[i] - lu.m[i][k] * b[k];
    }
    b[3] = b[3] / lu.m[3][3];
    for (int i = 2; i >= 0; --i) {
      float sum = 0.0f;
      for (int j = i + 1; j < 4; ++j) sum = sum + lu.m[i][j] * b[j];
      b[i] = (b[i] - sum) / lu.m[i][i];
    }
    for (int j = 0; j < 4; ++j) cols.m[r][j] = b[j];
  }
  return cols.transpose();
}

// ---------------------------------------------------------------------------------------------
// transform/transform.rs
struct Ray;
struct BBox;

struct Transform {
  M44 m, m_inv;
  Transform() = default;
  Transform(const M44& a, const M44& b) : m(a), m_inv(b) {}
  static Transform from_matrix(const M44& a) { return Transform(a, invert(a)); }  // :283-288
  Transform inverse() const { return Transform(m_inv, m); }                       // :33-39
  // :41-54
  static Transform translate(const V3& v) {
    return Transform(M44::rows(1, 0, 0, v.x, 0, 1, 0, v.y, 0, 0, 1, v.z, 0, 0, 0, 1),
                     M44::rows(1, 0, 0, -v.x, 0, 1, 0, -v.y, 0, 0, 1, -v.z, 0, 0, 0, 1));
  }
  // :56-69
  static Transform scale(float x, float y, float z) {
    return Transform(M44::rows(x, 0, 0, 0, 0, y, 0, 0, 0, 0, z, 0, 0, 0, 0, 1),
                     M44::rows(1.f / x, 0, 0, 0, 0, 1.f / y, 0, 0, 0, 0, 1.f / z, 0, 0, 0, 0, 1));
  }
  // :85-122
  static Transform rotate_x(float angle) {
    float s = std::sin(as_radians(angle)), c = std::cos(as_radians(angle));
    M44 a = M44::rows(1, 0, 0, 0, 0, c, -s, 0, 0, s, c, 0, 0, 0, 0, 1);
    return Transform(a, a.transpose());
  }
  static Transform rotate_y(float angle) {
    float s = std::sin(as_radians(angle)), c = std::cos(as_radians(angle));
    M44 a = M44::rows(c, 0, s, 0, 0, 1, 0, 0, -s, 0, c, 0, 0, 0, 0, 1);
    return Transform(a, a.transpose());
  }
  static Transform rotate_z(float angle) {
    float s = std::sin(as_radians(angle)), c = std::cos(as_radians(angle));
    M44 a = M44::rows(c, -s, 0, 0, s, c, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1);
    return Transform(a, a.transpose());
  }
  // :152-185.  Returns (m.invert(), m) where m is camera-to-world.
  static Transform look_at(const V3& pos, const V3& look, const V3& up) {
    M44 a;
    a.m[0][3] = pos.x;
    a.m[1][3] = pos.y;
    a.m[2][3] = pos.z;
    a.m[3][3] = 1.f;
    V3 dir = normalize(look - pos);
    V3 left = normalize(cross(normalize(up), dir));
    V3 new_up = cross(dir, left);
    a.m[0][0] = left.x;
    a.m[1][0] = left.y;
    a.m[2][0] = left.z;
    a.m[3][0] = 0.f;
    a.m[0][1] = new_up.x;
    a.m[1][1] = new_up.y;
    a.m[2][1] = new_up.z;
    a.m[3][1] = 0.f;
    a.m[0][2] = dir.x;
    a.m[1][2] = dir.y;
    a.m[2][2] = dir.z;
    a.m[3][2] = 0.f;
    return Transform(invert(a), a);
  }
  // :187-195
  bool swaps_handedness() const {
    return 0.f > (m.m[0][0] * (m.m[1][1] * m.m[2][2] - m.m[1][2] * m.m[2][1]) -
                  m.m[0][1] * (m.m[1][0] * m.m[2][2] - m.m[1][2] * m.m[2][0]) +
                  m.m[0][2] * (m.m[1][0] * m.m[2][1] - m.m[1][1] * m.m[2][0]));
  }
  // :207-219  Point: homogeneous divide when w != 1 (three true divides)
  V3 pt(const V3& p) const {
    float x = p.x, y = p.y, z = p.z;
    float xt = m.m[0][0] * x + m.m[0][1] * y + m.m[0][2] * z + m.m[0][3];
    float yt = m.m[1][0] * x + m.m[1][1] * y + m.m[1][2] * z + m.m[1][3];
    float zt = m.m[2][0] * x + m.m[2][1] * y + m.m[2][2] * z + m.m[2][3];
    float w = m.m[3][0] * x + m.m[3][1] * y + m.m[3][2] * z + m.m[3][3];
    if (w != 1.f) return V3(xt / w, yt / w, zt / w);
    return V3(xt, yt, zt);
  }
  // :221-229  Vector
  V3 vec(const V3& p) const {
    float x = p.x, y = p.y, z = p.z;
    return V3(m.m[0][0] * x + m.m[0][1] * y + m.m[0][2] * z,
              m.m[1][0] * x + m.m[1][1] * y + m.m[1][2] * z,
              m.m[2][0] * x + m.m[2][1] * y + m.m[2][2] * z);
  }
  // :231-239  Normal: transpose of the inverse
  V3 nrm(const V3& n) const {
    float x = n.x, y = n.y, z = n.z;
    return V3(m_inv.m[0][0] * x + m_inv.m[1][0] * y + m_inv.m[2][0] * z,
              m_inv.m[0][1] * x + m_inv.m[1][1] * y + m_inv.m[2][1] * z,
              m_inv.m[0][2] * x + m_inv.m[1][2] * y + m_inv.m[2][2] * z);
  }
};
// transform.rs:276-281:  (a*b).m = a.m*b.m ; (a*b).m_inv = b.m_inv*a.m_inv
inline Transform operator*(const Transform& a, const Transform& b) {
  return Transform(a.m * b.m, b.m_inv * a.m_inv);
}

// ---------------------------------------------------------------------------------------------
// ray.rs:7-60.  mint/maxt are interior-mutable in the reference (RefCell); here `mutable`.
struct Ray {
  V3 o, d;
  float time = 0.f;
  uint32_t depth = 0;
  mutable float mint = 0.f;
  mutable float maxt = F32_MAX;
  Ray() = default;
  Ray(const V3& o_, const V3& d_, float start) : o(o_), d(d_), mint(start) {}
  V3 at(float t) const { return o + (d * t); }  // ray.rs:59
};
// ray.rs:62-112
struct RayDifferential {
  Ray ray;
  bool has_differentials = false;
  V3 rx_origin, ry_origin, rx_dir, ry_dir;
  void scale_differentials(float s) {
    rx_origin = ray.o + (rx_origin - ray.o) * s;
    ry_origin = ray.o + (ry_origin - ray.o) * s;
    rx_dir = ray.d + (rx_dir - ray.d) * s;
    ry_dir = ray.d + (ry_dir - ray.d) * s;
  }
};
// transform.rs:241-248: only o and d move; mint/maxt/time/depth are copied.
inline Ray xf_ray(const Transform& t, const Ray& r) {
  Ray ret = r;
  ret.o = t.pt(r.o);
  ret.d = t.vec(r.d);
  return ret;
}

// ---------------------------------------------------------------------------------------------
// bbox.rs
struct BBox {
  V3 p_min{F32_MAX, F32_MAX, F32_MAX};
  V3 p_max{-F32_MAX, -F32_MAX, -F32_MAX};
  BBox() = default;
  BBox(const V3& a, const V3& b) : p_min(a), p_max(b) {}
  static BBox from_point(const V3& p) { return BBox(p, p); }
  // :158-170
  BBox united(const V3& pt) const {
    return BBox(V3(rmin(p_min.x, pt.x), rmin(p_min.y, pt.y), rmin(p_min.z, pt.z)),
                V3(rmax(p_max.x, pt.x), rmax(p_max.y, pt.y), rmax(p_max.z, pt.z)));
  }
  // :172-184
  BBox united(const BBox& b) const {
    return BBox(V3(rmin(p_min.x, b.p_min.x), rmin(p_min.y, b.p_min.y), rmin(p_min.z, b.p_min.z)),
                V3(rmax(p_max.x, b.p_max.x), rmax(p_max.y, b.p_max.y), rmax(p_max.z, b.p_max.z)));
  }
  // :51-54
  bool empty() const {
    V3 d = p_max - p_min;
    return d.x <= 0.f || d.y <= 0.f || d.z <= 0.f;
  }
  // :56-61
  bool overlaps(const BBox& b) const {
    bool x = p_max.x >= b.p_min.x && p_min.x <= b.p_max.x;
    bool y = p_max.y >= b.p_min.y && p_min.y <= b.p_max.y;
    bool z = p_max.z >= b.p_min.z && p_min.z <= b.p_max.z;
    return x && y && z;
  }
  // :63-67
  bool inside(const V3& p) const {
    return p.x >= p_min.x && p.x <= p_max.x && p.y >= p_min.y && p.y <= p_max.y &&
           p.z >= p_min.z && p.z <= p_max.z;
  }
  // :74-79
  float surface_area() const {
    float dx = rmax(p_max.x - p_min.x, 0.f);
    float dy = rmax(p_max.y - p_min.y, 0.f);
    float dz = rmax(p_max.z - p_min.z, 0.f);
    return 2.f * (dx * dy + dx * dz + dy * dz);
  }
  // :81-86
  float volume() const {
    float dx = rmax(p_max.x - p_min.x, 0.f);
    float dy = rmax(p_max.y - p_min.y, 0.f);
    float dz = rmax(p_max.z - p_min.z, 0.f);
    return dx * dy * dz;
  }
  // :88-97 (tie rules: x only if strictly largest; else y if strictly > z; else z)
  int max_extent() const {
    V3 d = p_max - p_min;
    if (d.x > d.y && d.x > d.z) return 0;
    if (d.y > d.z) return 1;
    return 2;
  }
  // :185-209  slab test, starting from the ray's LIVE (mint, maxt); per-axis true divide.
  bool intersect(const Ray& r, float* t0_out = nullptr, float* t1_out = nullptr) const {
    float t0 = r.mint, t1 = r.maxt;
    for (int i = 0; i < 3; ++i) {
      float inv_ray_dir = 1.f / r.d[i];
      float t_a = (p_min[i] - r.o[i]) * inv_ray_dir;
      float t_b = (p_max[i] - r.o[i]) * inv_ray_dir;
      if (t_a > t_b) std::swap(t_a, t_b);
      t0 = rmax(t_a, t0);
      t1 = rmin(t_b, t1);
      if (t0 > t1) return false;
    }
    if (t0_out) *t0_out = t0;
    if (t1_out) *t1_out = t1;
    return true;
  }
};
// transform.rs:256-273
inline BBox xf_bbox(const Transform& t, const BBox& b) {
  V3 tx = t.vec(V3(b.p_max.x - b.p_min.x, 0.f, 0.f));
  V3 ty = t.vec(V3(0.f, b.p_max.y - b.p_min.y, 0.f));
  V3 tz = t.vec(V3(0.f, 0.f, b.p_max.z - b.p_min.z));
  V3 tp = t.pt(b.p_min);
  return BBox::from_point(tp)
      .united(tp + tx)
      .united(tp + ty)
      .united(tp + tz)
      .united(tp + tx + ty)
      .united(tp + tx + tz)
      .united(tp + ty + tz)
      .united(tp + tx + ty + tz);
}

// ---------------------------------------------------------------------------------------------
// utils/mod.rs:56-92
inline bool quadratic(float a, float b, float c, float* t0, float* t1) {
  float descrim = b * b - 4.f * a * c;
  if (descrim < 0.0f) return false;
  if (std::fabs(descrim) < 1e-6f) {
    if (a == 0.0f) return false;
    float t = -b / (2.0f * a);
    *t0 = t;
    *t1 = t;
    return true;
  }
  float root_descrim = std::sqrt(descrim);
  float q = (b < 0.0f) ? -0.5f * (b - root_descrim) : -0.5f * (b + root_descrim);
  if (a == 0.0f) return false;
  float x0 = q / a;
  float x1 = c / q;
  if (x0 < x1) {
    *t0 = x0;
    *t1 = x1;
  } else {
    *t0 = x1;
    *t1 = x0;
  }
  return true;
}

// utils/mod.rs:94-110
inline bool solve_linear_system_2x2(const float a[2][2], const float b[2], float* x0, float* x1) {
  float det = a[0][0] * a[1][1] - a[0][1] * a[1][0];
  if (std::fabs(det) < 1e-10f) return false;
  float inv_det = 1.0f / det;
  float r0 = (a[1][1] * b[0] - a[0][1] * b[1]) * inv_det;
  float r1 = (a[0][0] * b[1] - a[1][0] * b[0]) * inv_det;
  if (std::isnan(r0) || std::isnan(r1)) return false;
  *x0 = r0;
  *x1 = r1;
  return true;
}

// utils/mod.rs:112-169  quick-select style partition; restated iteratively (the reference's
// tail calls only ever recurse into one side).  `key(i)` reads element i, `swp(i,j)` swaps.
template <class Key, class Swap>
inline void partition_by(size_t lo, size_t n, Key key, Swap swp) {
  for (;;) {
    if (n < 3) {
      if (n == 2 && key(lo + 1) < key(lo)) swp(lo, lo + 1);
      return;
    }
    auto fst = key(lo);
    auto mid = key(lo + n / 2);
    auto lst = key(lo + n - 1);
    auto pivot = (fst < mid && mid < lst) ? mid : ((mid < fst && fst < lst) ? fst : lst);
    size_t last_smaller = 0, num_pivots = 0;
    for (size_t i = 0; i < n; ++i) {
      auto bv = key(lo + i);
      if (bv < pivot) {
        swp(lo + last_smaller + num_pivots, lo + i);
        swp(lo + last_smaller + num_pivots, lo + last_smaller);
        last_smaller += 1;
      } else if (bv == pivot) {
        swp(lo + last_smaller + num_pivots, lo + i);
        num_pivots += 1;
      }
    }
    size_t pivot_idx = last_smaller > 1 ? last_smaller : 1;
    if (pivot_idx + num_pivots <= n / 2) {
      lo = lo + pivot_idx;
      n = n - pivot_idx;
    } else if (pivot_idx >= n / 2) {
      n = pivot_idx;
    } else {
      return;
    }
  }
}

// utils/mod.rs:171-205
inline void get_num_subwindows_2d(size_t count, size_t aspect, size_t* nx, size_t* ny) {
  size_t x = 1, y = count;
  while ((y % 2) == 0 && 2 * aspect * x < y) {
    y /= 2;
    x *= 2;
  }
  *nx = x;
  *ny = y;
}
inline void get_crop_window(size_t num, size_t count, float aspect, float out[4]) {
  size_t nx, ny;
  if (aspect < 1.0f) {
    size_t inva = f2usize(1.0f / aspect);
    get_num_subwindows_2d(count, inva, &nx, &ny);
  } else {
    size_t x, y;
    get_num_subwindows_2d(count, f2usize(aspect), &x, &y);
    nx = y;
    ny = x;
  }
  size_t xo = num % nx, yo = num / nx;
  out[0] = (float)xo / (float)nx;
  out[1] = (float)(xo + 1) / (float)nx;
  out[2] = (float)yo / (float)ny;
  out[3] = (float)(yo + 1) / (float)ny;
}

}  // namespace orc
