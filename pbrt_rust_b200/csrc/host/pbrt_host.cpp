// Host-side mirror of the pbrt_rust constructors above the drop-in boundary (include/
// pbrtb200_host.h) and the flatten shim that turns them into a pbrtb200_scene.  Plain C++17, no
// CUDA, never touches the CPU oracle.  Float arithmetic keeps the reference's operation order
// (compile with -ffp-contract=off) because the BVH it builds must be the reference's tree.
#include <algorithm>
#include <zlib.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/pbrtb200_host.h"
#include "../host_logic.hpp"

namespace {

constexpr float kF32Max = 3.402823466e+38f;
constexpr float kPi = 3.14159265358979323846f;

typedef float Mat[16];  // row-major 4x4

struct Vec {
  float x, y, z;
};
inline Vec operator-(Vec a, Vec b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec operator+(Vec a, Vec b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec cross(Vec a, Vec v) {
  return {(a.y * v.z) - (a.z * v.y), (a.z * v.x) - (a.x * v.z), (a.x * v.y) - (a.y * v.x)};
}
inline Vec normalize(Vec a) {  // normal.rs:180-185 : v * (1/len)
  const float l = std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z);
  const float r = 1.0f / l;
  return {a.x * r, a.y * r, a.z * r};
}
inline float radians(float d) { return d * kPi / 180.0f; }  // utils/mod.rs:36-39

void mat_identity(Mat m) {
  for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.f : 0.f;
}
void mat_mul(const Mat a, const Mat b, Mat r) {  // matrix4x4.rs:167-181
  Mat t;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      t[4 * i + j] = a[4 * i] * b[j] + a[4 * i + 1] * b[4 + j] + a[4 * i + 2] * b[8 + j] +
                     a[4 * i + 3] * b[12 + j];
  std::memcpy(r, t, sizeof t);
}
void mat_transpose(const Mat a, Mat r) {
  Mat t;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) t[4 * i + j] = a[4 * j + i];
  std::memcpy(r, t, sizeof t);
}
// matrix4x4.rs:66-165: LU with scaled partial pivoting + four solves; false = "Singular matrix!"
bool mat_invert(const Mat a, Mat out) {
  float s[4];
  for (int i = 0; i < 4; ++i) {
    float acc = 0.f;
    for (int j = 0; j < 4; ++j) acc = std::fmax(std::fabs(a[4 * i + j]), acc);
    s[i] = acc;
  }
  Mat lu;
  std::memcpy(lu, a, sizeof lu);
  int piv[4] = {0, 1, 2, 3};
  for (int k = 0; k < 3; ++k) {
    float c = 0.f;
    int p = k;
    for (int i = k; i < 4; ++i) {
      const float cc = std::fabs(lu[4 * i + k] / s[i]);
      if (cc > c) {
        c = cc;
        p = i;
      }
    }
    piv[k] = p;
    if (c == 0.f) return false;
    if (p != k)
      for (int j = k; j < 4; ++j) std::swap(lu[4 * k + j], lu[4 * p + j]);
    for (int i = k + 1; i < 4; ++i) {
      const float mi = lu[4 * i + k] / lu[4 * k + k];
      lu[4 * i + k] = mi;
      for (int j = k + 1; j < 4; ++j) lu[4 * i + j] = lu[4 * i + j] - mi * lu[4 * k + j];
    }
  }
  if (std::fabs(lu[15]) < 1.0e-6f) return false;
  for (int r = 0; r < 4; ++r) {
    float b[4] = {0.f, 0.f, 0.f, 0.f};
    b[r] = 1.f;
    for (int k = 0; k < 3; ++k) {
      if (piv[k] != k) std::swap(b[piv[k]], b[k]);
      for (int i = k + 1; i < 4; ++i) b[i] = b[i] - lu[4 * i + k] * b[k];
    }
    b[3] = b[3] / lu[15];
    for (int i = 2; i >= 0; --i) {
      float sum = 0.f;
      for (int j = i + 1; j < 4; ++j) sum = sum + lu[4 * i + j] * b[j];
      b[i] = (b[i] - sum) / lu[4 * i + i];
    }
    for (int j = 0; j < 4; ++j) out[4 * j + r] = b[j];  // transpose of the row-wise solutions
  }
  return true;
}
// transform.rs:207-229
Vec xf_point(const Mat m, Vec p) {
  const float xt = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3];
  const float yt = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7];
  const float zt = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11];
  const float w = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15];
  if (w != 1.f) return {xt / w, yt / w, zt / w};
  return {xt, yt, zt};
}
Vec xf_vector(const Mat m, Vec v) {
  return {m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z,
          m[8] * v.x + m[9] * v.y + m[10] * v.z};
}
bool swaps_handedness(const Mat m) {  // transform.rs:187-195
  return 0.f > (m[0] * (m[5] * m[10] - m[6] * m[9]) - m[1] * (m[4] * m[10] - m[6] * m[8]) +
                m[2] * (m[4] * m[9] - m[5] * m[8]));
}

struct Box {
  float lo[3] = {kF32Max, kF32Max, kF32Max};
  float hi[3] = {-kF32Max, -kF32Max, -kF32Max};
  void grow(Vec p) {  // bbox.rs:158-170
    lo[0] = std::fmin(lo[0], p.x);
    lo[1] = std::fmin(lo[1], p.y);
    lo[2] = std::fmin(lo[2], p.z);
    hi[0] = std::fmax(hi[0], p.x);
    hi[1] = std::fmax(hi[1], p.y);
    hi[2] = std::fmax(hi[2], p.z);
  }
  void grow(const Box& b) {  // bbox.rs:172-184
    for (int i = 0; i < 3; ++i) {
      lo[i] = std::fmin(lo[i], b.lo[i]);
      hi[i] = std::fmax(hi[i], b.hi[i]);
    }
  }
  float area() const {  // bbox.rs:74-79
    const float dx = std::fmax(hi[0] - lo[0], 0.f), dy = std::fmax(hi[1] - lo[1], 0.f),
                dz = std::fmax(hi[2] - lo[2], 0.f);
    return 2.f * (dx * dy + dx * dz + dy * dz);
  }
  int max_extent() const {  // bbox.rs:88-97
    const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    if (dx > dy && dx > dz) return 0;
    if (dy > dz) return 1;
    return 2;
  }
};

struct MeshRec {
  Mat o2w, o2w_inv;
  bool ro;
  std::vector<uint32_t> vi;
  std::vector<Vec> p;  // world space (Mesh::new, mesh.rs:308)
  std::vector<float> n, s, uv;
  int material, area_light;
};
struct SphereRec {  // a quadric: sphere, cylinder or disk (include/pbrtb200.h)
  Mat o2w, o2w_inv;
  bool ro;
  float radius, z_min, z_max, phi_max, theta_min, theta_max;
  int material;
  uint32_t kind = PBRTB200_QUADRIC_SPHERE;
};
struct Object {
  int kind;  // 0 mesh, 1 sphere
  uint32_t index;
};
// One refined primitive handed to the BVH builder.
struct PrimRef {
  uint32_t kind;    // 0 tri, 1 sphere
  uint32_t object;  // object ordinal
  uint32_t index;   // triple index / 0
};

}  // namespace

struct pbh_scene {
  std::string err;
  std::vector<pbrtb200_texture> textures;
  std::vector<pbrtb200_material> materials;
  std::vector<pbrtb200_light> lights;
  std::vector<MeshRec> meshes;
  std::vector<SphereRec> spheres;
  std::vector<Object> objects;
  // build output
  std::vector<pbrtb200_node32> nodes;
  std::vector<PrimRef> ordered;
  // flattened arrays
  std::vector<uint32_t> leaf_prim;
  std::vector<pbrtb200_tri48> tris;
  std::vector<pbrtb200_sphere80> fspheres;
  std::vector<float> sphere_o2w;
  std::vector<pbrtb200_mesh> fmeshes;
  std::vector<float> tri_uv, tri_n, tri_s;
  std::vector<uint32_t> area_prims;
  std::vector<pbrtb200_mipmap> mipmaps;
  std::vector<float> texels;  // 4 floats per texel, all levels of all mipmaps
  pbrtb200_scene flat{};
  bool built = false;
};

namespace {

// ---- the in-place BVH builder ------------------------------------------------------------------
// Works on one index permutation `idx` over the refined primitives.  Reproduces, node for node,
// recursive_build (bvh.rs:189-258): `into_iter().partition` is a STABLE partition; partition_by
// (utils/mod.rs:112-169) is replayed swap for swap; nodes are emitted in the depth-first order of
// PackedBVHNode::flatten_tree (bvh.rs:282-316), which also makes leaf prim offsets a running count.
struct Builder {
  const std::vector<Box>& bounds;
  const std::vector<Vec>& centroid;
  std::vector<uint32_t>& idx;
  std::vector<uint32_t> tmp;
  std::vector<pbrtb200_node32>& out;
  size_t max_prims;
  int method;  // 0 middle, 1 equal, 2 sah
  std::string* err;
  static constexpr size_t kParallelMin = 1u << 16;  // primitives in a subtree worth a thread
  static constexpr int kParallelDepth = 5;           // up to 2^5 concurrent subtree builders

  float key(size_t i, int dim) const {
    const Vec& c = centroid[idx[i]];
    return dim == 0 ? c.x : (dim == 1 ? c.y : c.z);
  }

  void select_equal(size_t lo, size_t n, int dim) {
    for (;;) {
      if (n < 3) {
        if (n == 2 && key(lo + 1, dim) < key(lo, dim)) std::swap(idx[lo], idx[lo + 1]);
        return;
      }
      const float fst = key(lo, dim), mid = key(lo + n / 2, dim), lst = key(lo + n - 1, dim);
      const float pivot = (fst < mid && mid < lst) ? mid : ((mid < fst && fst < lst) ? fst : lst);
      size_t smaller = 0, pivots = 0;
      for (size_t i = 0; i < n; ++i) {
        const float bv = key(lo + i, dim);
        if (bv < pivot) {
          std::swap(idx[lo + smaller + pivots], idx[lo + i]);
          std::swap(idx[lo + smaller + pivots], idx[lo + smaller]);
          ++smaller;
        } else if (bv == pivot) {
          std::swap(idx[lo + smaller + pivots], idx[lo + i]);
          ++pivots;
        }
      }
      const size_t pi = std::max<size_t>(smaller, 1);
      if (pi + pivots <= n / 2) {
        lo += pi;
        n -= pi;
      } else if (pi >= n / 2) {
        n = pi;
      } else {
        return;
      }
    }
  }

  template <class Pred>
  size_t stable_split(size_t lo, size_t hi, Pred left) {
    size_t a = lo, t = 0;
    if (tmp.size() < hi - lo) tmp.resize(hi - lo);
    for (size_t i = lo; i < hi; ++i) {
      const uint32_t v = idx[i];
      if (left(v))
        idx[a++] = v;
      else
        tmp[t++] = v;
    }
    std::copy(tmp.begin(), tmp.begin() + (ptrdiff_t)t, idx.begin() + (ptrdiff_t)a);
    return a;
  }

  void emit_leaf(const Box& b, size_t lo, size_t hi) {
    pbrtb200_node32 nd{};
    std::memcpy(nd.bmin, b.lo, 12);
    std::memcpy(nd.bmax, b.hi, 12);
    nd.offset = (uint32_t)lo;
    nd.count = (uint16_t)std::min<size_t>(hi - lo, 65535);
    if (hi - lo > 65535 && err->empty()) *err = "leaf with more than 65535 primitives";
    nd.is_leaf = 1;
    out.push_back(nd);
  }

  // returns false on the reference's assert!(p.len() > 0) (bvh.rs:237-238)
  bool build(size_t lo, size_t hi, int depth = 0) {
    Box bbox;
    for (size_t i = lo; i < hi; ++i) bbox.grow(bounds[idx[i]]);
    const size_t n = hi - lo;
    if (n == 1) {
      emit_leaf(bbox, lo, hi);
      return true;
    }
    Box cb;
    for (size_t i = lo; i < hi; ++i) cb.grow(centroid[idx[i]]);
    const int dim = cb.max_extent();
    if (cb.lo[dim] == cb.hi[dim]) {
      emit_leaf(bbox, lo, hi);
      return true;
    }
    size_t mid;
    if (method == 0) {  // split_middle, bvh.rs:86-90
      const float p_mid = 0.5f * (cb.lo[dim] + cb.hi[dim]);
      mid = stable_split(lo, hi, [&](uint32_t v) {
        const Vec& c = centroid[v];
        return (dim == 0 ? c.x : (dim == 1 ? c.y : c.z)) < p_mid;
      });
    } else if (method == 1 || n <= 4) {  // split_equal_counts, bvh.rs:92-104 (SAH: n <= 4)
      select_equal(lo, n, dim);
      mid = lo + n / 2;
    } else {  // split_surface_area_heuristic, bvh.rs:106-186
      constexpr int NB = 12;
      size_t cnt[NB] = {0};
      Box bb[NB];
      const float cmin = cb.lo[dim], cmax = cb.hi[dim];
      auto bucket = [&](uint32_t v) -> size_t {
        const Vec& c = centroid[v];
        const float pdist = (dim == 0 ? c.x : (dim == 1 ? c.y : c.z)) - cmin;
        const float dist = cmax - cmin;
        size_t b = (size_t)pbh::sat_usize((float)NB * (pdist / dist));
        return b == (size_t)NB ? (size_t)NB - 1 : b;
      };
      for (size_t i = lo; i < hi; ++i) {
        const size_t b = bucket(idx[i]);
        if (b >= (size_t)NB) {
          *err = "SAH bucket index out of bounds";
          return false;
        }
        cnt[b]++;
        bb[b].grow(bounds[idx[i]]);
      }
      const float tsa = bbox.area();
      size_t best = 0;
      float best_cost = kF32Max;
      for (int i = 0; i < NB - 1; ++i) {
        size_t c0 = 0, c1 = 0;
        Box b0, b1;
        for (int j = 0; j <= i; ++j) {
          c0 += cnt[j];
          b0.grow(bb[j]);
        }
        for (int j = i + 1; j < NB; ++j) {
          c1 += cnt[j];
          b1.grow(bb[j]);
        }
        const float cost = 0.125f * ((float)c0 * b0.area() + (float)c1 * b1.area()) / tsa;
        if (cost < best_cost) {
          best = (size_t)i;
          best_cost = cost;
        }
      }
      if (max_prims < n || pbh::sat_usize(best_cost) < n) {
        mid = stable_split(lo, hi, [&](uint32_t v) { return bucket(v) <= best; });
      } else {
        emit_leaf(bbox, lo, hi);
        return true;
      }
    }
    if (mid == lo || mid == hi) {
      *err = "BVH split produced an empty side (assert!(p.len() > 0), bvh.rs:237-238)";
      return false;
    }
    const size_t me = out.size();
    out.push_back(pbrtb200_node32{});
    uint32_t second;
    if (n >= kParallelMin && depth < kParallelDepth) {
      // Large subtrees: build the two children concurrently, each into its own node vector
      // (inner-node child indices are relative to that vector), then splice them in depth-first
      // order.  The ranges of `idx` are disjoint, so the result is the sequential builder's tree.
      std::vector<pbrtb200_node32> lv, rv;
      std::string lerr, rerr;
      Builder lb{bounds, centroid, idx, {}, lv, max_prims, method, &lerr};
      Builder rb{bounds, centroid, idx, {}, rv, max_prims, method, &rerr};
      bool lok = false, rok = false;
      std::thread th([&] { lok = lb.build(lo, mid, depth + 1); });
      rok = rb.build(mid, hi, depth + 1);
      th.join();
      if (!lok || !rok || !lerr.empty() || !rerr.empty()) {
        *err = !lerr.empty() ? lerr : rerr;
        return false;
      }
      auto splice = [&](const std::vector<pbrtb200_node32>& v) {
        const uint32_t base = (uint32_t)out.size();
        for (pbrtb200_node32 nd : v) {
          if (!nd.is_leaf) nd.offset += base;
          out.push_back(nd);
        }
        return base;
      };
      splice(lv);
      second = splice(rv);
    } else {
      if (!build(lo, mid, depth + 1)) return false;
      second = (uint32_t)out.size();
      if (!build(mid, hi, depth + 1)) return false;
    }
    // bvh.rs:244-245: bounds = left.bounds U right.bounds
    const pbrtb200_node32 &l = out[me + 1], &r = out[second];
    pbrtb200_node32& nd = out[me];
    for (int i = 0; i < 3; ++i) {
      nd.bmin[i] = std::fmin(l.bmin[i], r.bmin[i]);
      nd.bmax[i] = std::fmax(l.bmax[i], r.bmax[i]);
    }
    nd.offset = second;
    nd.count = 0;
    nd.axis = (uint8_t)dim;
    nd.is_leaf = 0;
    return true;
  }
};

Box sphere_world_bound(const SphereRec& s) {  // sphere.rs:112-128 + transform.rs:256-273
  const Vec lo = {-s.radius, -s.radius, s.z_min}, hi = {s.radius, s.radius, s.z_max};
  const Vec tx = xf_vector(s.o2w, {hi.x - lo.x, 0.f, 0.f});
  const Vec ty = xf_vector(s.o2w, {0.f, hi.y - lo.y, 0.f});
  const Vec tz = xf_vector(s.o2w, {0.f, 0.f, hi.z - lo.z});
  const Vec tp = xf_point(s.o2w, lo);
  Box b;
  b.lo[0] = b.hi[0] = tp.x;
  b.lo[1] = b.hi[1] = tp.y;
  b.lo[2] = b.hi[2] = tp.z;
  b.grow(tp + tx);
  b.grow(tp + ty);
  b.grow(tp + tz);
  b.grow(tp + tx + ty);
  b.grow(tp + tx + tz);
  b.grow(tp + ty + tz);
  b.grow(tp + tx + ty + tz);
  return b;
}

void rows3(const Mat m, float out[12]) { std::memcpy(out, m, 48); }

float filter_eval(int ty, float xw, float yw, float p0, float p1, float x, float y) {
  const float inv_xw = 1.0f / xw, inv_yw = 1.0f / yw;  // filter.rs:12-19
  switch (ty) {
    case 0:
      return 1.0f;
    case 1: {
      const float dx = (xw - std::fabs(x)) * inv_xw, dy = (yw - std::fabs(y)) * inv_yw;
      return std::fmax(dx, 0.0f) * std::fmax(dy, 0.0f);
    }
    case 2: {
      const float ex = std::exp(-p0 * xw * xw), ey = std::exp(-p0 * yw * yw);
      auto g = [&](float v, float e) { return std::fmax(std::exp(-p0 * v * v) - e, 0.0f); };
      return g(x, ex) * g(y, ey);
    }
    case 3: {
      const float b = p0, c = p1;
      auto mitchell = [&](float v) {
        const float t = std::fabs(v * 2.0f);
        float r;
        if (t >= 2.0f)
          r = 0.0f;
        else if (t > 1.0f)
          r = (-b - 6.0f * c) * t * t * t + (6.0f * b + 30.0f * c) * t * t +
              (-12.0f * b - 48.0f * c) * t + (8.0f * b + 24.0f * c);
        else
          r = (12.0f - 9.0f * b - 6.0f * c) * t * t * t + (-18.0f + 12.0f * b + 6.0f * c) * t * t +
              (6.0f - 2.0f * b);
        return (1.0f / 6.0f) * r;
      };
      return mitchell(x * inv_xw) * mitchell(y * inv_yw);
    }
    default: {
      auto sinc = [&](float xx) {  // utils/mod.rs:207-218
        float v = std::fabs(xx);
        if (v < 1e-5f) return 1.0f;
        if (v >= 1.0f) return 0.0f;
        v *= kPi;
        const float vtau = v * p0;
        const float s = std::sin(vtau) / vtau;
        return s * std::sin(v) / v;
      };
      return sinc(x * inv_xw) * sinc(y * inv_yw);
    }
  }
}

}  // namespace

extern "C" {

void pbh_translate(const float v[3], float m[16], float minv[16]) {
  mat_identity(m);
  mat_identity(minv);
  m[3] = v[0];
  m[7] = v[1];
  m[11] = v[2];
  minv[3] = -v[0];
  minv[7] = -v[1];
  minv[11] = -v[2];
}
void pbh_scale(float x, float y, float z, float m[16], float minv[16]) {
  mat_identity(m);
  mat_identity(minv);
  m[0] = x;
  m[5] = y;
  m[10] = z;
  minv[0] = 1.f / x;
  minv[5] = 1.f / y;
  minv[10] = 1.f / z;
}
void pbh_rotate_x(float deg, float m[16], float minv[16]) {
  const float s = std::sin(radians(deg)), c = std::cos(radians(deg));
  mat_identity(m);
  m[5] = c;
  m[6] = -s;
  m[9] = s;
  m[10] = c;
  mat_transpose(m, minv);
}
void pbh_rotate_y(float deg, float m[16], float minv[16]) {
  const float s = std::sin(radians(deg)), c = std::cos(radians(deg));
  mat_identity(m);
  m[0] = c;
  m[2] = s;
  m[8] = -s;
  m[10] = c;
  mat_transpose(m, minv);
}
void pbh_rotate_z(float deg, float m[16], float minv[16]) {
  const float s = std::sin(radians(deg)), c = std::cos(radians(deg));
  mat_identity(m);
  m[0] = c;
  m[1] = -s;
  m[4] = s;
  m[5] = c;
  mat_transpose(m, minv);
}
void pbh_mul(const float am[16], const float aminv[16], const float bm[16], const float bminv[16],
             float m[16], float minv[16]) {  // transform.rs:276-281
  Mat r, ri;
  mat_mul(am, bm, r);
  mat_mul(bminv, aminv, ri);
  std::memcpy(m, r, sizeof r);
  std::memcpy(minv, ri, sizeof ri);
}
int pbh_invert(const float m[16], float out[16]) {
  Mat r;
  if (!mat_invert(m, r)) return PBRTB200_ESINGULAR;
  std::memcpy(out, r, sizeof r);
  return PBRTB200_OK;
}
int pbh_look_at(const float pos[3], const float look[3], const float up[3], float m[16],
                float minv[16]) {  // transform.rs:152-185: returns (c2w.invert(), c2w)
  Mat c2w;
  mat_identity(c2w);
  c2w[3] = pos[0];
  c2w[7] = pos[1];
  c2w[11] = pos[2];
  c2w[15] = 1.f;
  const Vec dir = normalize(Vec{look[0], look[1], look[2]} - Vec{pos[0], pos[1], pos[2]});
  const Vec left = normalize(cross(normalize(Vec{up[0], up[1], up[2]}), dir));
  const Vec nup = cross(dir, left);
  c2w[0] = left.x;
  c2w[4] = left.y;
  c2w[8] = left.z;
  c2w[12] = 0.f;
  c2w[1] = nup.x;
  c2w[5] = nup.y;
  c2w[9] = nup.z;
  c2w[13] = 0.f;
  c2w[2] = dir.x;
  c2w[6] = dir.y;
  c2w[10] = dir.z;
  c2w[14] = 0.f;
  Mat inv;
  if (!mat_invert(c2w, inv)) return PBRTB200_ESINGULAR;
  std::memcpy(m, inv, sizeof inv);
  std::memcpy(minv, c2w, sizeof c2w);
  return PBRTB200_OK;
}

pbh_scene* pbh_scene_new(void) { return new pbh_scene(); }
void pbh_scene_free(pbh_scene* s) { delete s; }
const char* pbh_last_error(const pbh_scene* s) { return s ? s->err.c_str() : ""; }

int pbh_texture_constant(pbh_scene* s, const float rgb[3]) {
  pbrtb200_texture t{};
  t.kind = PBRTB200_TEX_CONSTANT;
  std::memcpy(t.value, rgb, 12);
  s->textures.push_back(t);
  return (int)s->textures.size() - 1;
}
int pbh_texture_checkerboard(pbh_scene* s, int map_kind, const float map[12], int tex1, int tex2,
                             int antialiased) {
  pbrtb200_texture t{};
  t.kind = PBRTB200_TEX_CHECKER2D;
  t.map_kind = map_kind;
  std::memcpy(t.map, map, 48);
  t.tex1 = tex1;
  t.tex2 = tex2;
  t.aa = antialiased ? 1 : 0;
  s->textures.push_back(t);
  return (int)s->textures.size() - 1;
}
int pbh_texture_uv(pbh_scene* s, int map_kind, const float map[12]) {
  pbrtb200_texture t{};
  t.kind = PBRTB200_TEX_UV;
  t.map_kind = map_kind;
  std::memcpy(t.map, map, 48);
  s->textures.push_back(t);
  return (int)s->textures.size() - 1;
}
int pbh_mapping_from_transform(const float m[16], float map[12]) {
  if (m[12] != 0.f || m[13] != 0.f || m[14] != 0.f || m[15] != 1.f) return PBRTB200_EINVAL;
  std::memcpy(map, m, 48);
  return PBRTB200_OK;
}
int pbh_texture_scale(pbh_scene* s, int tex1, int tex2) {
  pbrtb200_texture t{};
  t.kind = PBRTB200_TEX_SCALE;
  t.tex1 = tex1;
  t.tex2 = tex2;
  s->textures.push_back(t);
  return (int)s->textures.size() - 1;
}
int pbh_texture_mix(pbh_scene* s, int tex1, int tex2, int amount) {
  pbrtb200_texture t{};
  t.kind = PBRTB200_TEX_MIX;
  t.tex1 = tex1;
  t.tex2 = tex2;
  t.tex3 = amount;
  s->textures.push_back(t);
  return (int)s->textures.size() - 1;
}
int pbh_texture_bilerp(pbh_scene* s, int map_kind, const float map[12], const float v00[3],
                       const float v01[3], const float v10[3], const float v11[3]) {
  pbrtb200_texture t{};
  t.kind = PBRTB200_TEX_BILERP;
  t.map_kind = map_kind;
  std::memcpy(t.map, map, 48);
  std::memcpy(t.value + 0, v00, 12);
  std::memcpy(t.value + 3, v01, 12);
  std::memcpy(t.value + 6, v10, 12);
  std::memcpy(t.value + 9, v11, 12);
  s->textures.push_back(t);
  return (int)s->textures.size() - 1;
}
int pbh_texture_dots(pbh_scene* s, int map_kind, const float map[12], int inside, int outside) {
  pbrtb200_texture t{};
  t.kind = PBRTB200_TEX_DOTS;
  t.map_kind = map_kind;
  std::memcpy(t.map, map, 48);
  t.tex1 = inside;
  t.tex2 = outside;
  s->textures.push_back(t);
  return (int)s->textures.size() - 1;
}
static int noise_texture(pbh_scene* s, int kind, int octaves, float roughness, const float w2t[12]) {
  if (octaves < 0) {  // f32::clamp(0.0, max) panics when max < min (noise.rs:116)
    s->err = "fbm/wrinkled texture: octaves must be >= 0";
    return PBRTB200_EINVAL;
  }
  pbrtb200_texture t{};
  t.kind = kind;
  t.map_kind = PBRTB200_MAP_IDENTITY3D;
  std::memcpy(t.map, w2t, 48);
  t.value[0] = roughness;
  t.aa = octaves;
  s->textures.push_back(t);
  return (int)s->textures.size() - 1;
}
int pbh_texture_fbm(pbh_scene* s, int octaves, float roughness, const float w2t[12]) {
  return noise_texture(s, PBRTB200_TEX_FBM, octaves, roughness, w2t);
}
int pbh_texture_wrinkled(pbh_scene* s, int octaves, float roughness, const float w2t[12]) {
  return noise_texture(s, PBRTB200_TEX_WRINKLED, octaves, roughness, w2t);
}

// ---- ImageTexture / MIPMap construction (texture/mipmap.rs, texture/imagemap.rs) ----------------
namespace {
struct Tex3 {
  float r, g, b;
};
inline Tex3 operator+(Tex3 a, Tex3 b) { return {a.r + b.r, a.g + b.g, a.b + b.b}; }
inline Tex3 operator*(Tex3 a, float s) { return {a.r * s, a.g * s, a.b * s}; }

inline float mip_sinc(float x, float tau) {  // utils/mod.rs:207-217
  float v = std::fabs(x);
  if (v < 1e-5f) return 1.0f;
  if (v >= 1.0f) return 0.0f;
  v *= 3.14159265358979323846f;
  const float vtau = v * tau;
  const float s = std::sin(vtau) / vtau;
  return s * std::sin(v) / v;
}
inline int32_t mip_mod(int32_t a, int32_t b) {  // utils/mod.rs:219-223
  const int32_t x = a - (a / b) * b;
  return x < 0 ? x + b : x;
}
// texel index under an ImageWrap, or -1 (Black outside / resampling tap outside)
inline int32_t mip_wrap(int32_t i, int32_t dim, int wrap, bool black_is_raw) {
  if (wrap == PBRTB200_WRAP_REPEAT) return mip_mod(i, dim);
  if (wrap == PBRTB200_WRAP_CLAMP) return i < 0 ? 0 : (i > dim - 1 ? dim - 1 : i);
  (void)black_is_raw;
  return (i >= 0 && i < dim) ? i : -1;
}
struct Taps {
  int32_t first;
  float w[4];
};
std::vector<Taps> mip_resample_weights(size_t oldres, size_t newres) {  // mipmap.rs:22-43
  std::vector<Taps> out(newres);
  for (size_t i = 0; i < newres; ++i) {
    const float center = ((float)i + 0.5f) * (float)oldres / (float)newres;
    Taps t;
    t.first = pbh::sat_i32(std::floor((center - 2.0f) + 0.5f));
    for (int j = 0; j < 4; ++j) t.w[j] = mip_sinc((((float)(t.first + j) + 0.5f) - center) / 2.0f, 2.0f);
    float sum = 0.0f;
    for (int j = 0; j < 4; ++j) sum = sum + t.w[j];
    const float inv = 1.0f / sum;
    for (int j = 0; j < 4; ++j) t.w[j] *= inv;
    out[i] = t;
  }
  return out;
}
// mipmap.rs:45-106.  The t pass is in place, like the reference: a row that was already
// resampled is read back by the rows after it.
void mip_resize_pot(size_t w, size_t h, const std::vector<Tex3>& px, int wrap, size_t* wp, size_t* hp,
                    std::vector<Tex3>* out) {
  size_t wpot = 1, hpot = 1;
  while (wpot < w) wpot <<= 1;
  while (hpot < h) hpot <<= 1;
  std::vector<Tex3> np(wpot * hpot);
  const auto sw = mip_resample_weights(w, wpot);
  for (size_t t = 0; t < h; ++t)
    for (size_t s = 0; s < wpot; ++s) {
      Tex3 acc{0.f, 0.f, 0.f};
      for (int j = 0; j < 4; ++j) {
        const int32_t o = mip_wrap(sw[s].first + j, (int32_t)w, wrap, true);
        if (o >= 0) acc = px[t * w + (size_t)o] * sw[s].w[j] + acc;
      }
      np[t * wpot + s] = acc;
    }
  for (size_t t = h; t < hpot; ++t)
    for (size_t s = 0; s < wpot; ++s) np[t * wpot + s] = px[0];
  const auto tw = mip_resample_weights(h, hpot);
  for (size_t s = 0; s < wpot; ++s)
    for (size_t t = 0; t < hpot; ++t) {
      Tex3 acc{0.f, 0.f, 0.f};
      for (int j = 0; j < 4; ++j) {
        const int32_t o = mip_wrap(tw[t].first + j, (int32_t)h, wrap, true);
        if (o >= 0) acc = np[(size_t)o * wpot + s] * tw[t].w[j] + acc;
      }
      np[t * wpot + s] = acc;
    }
  *wp = wpot;
  *hp = hpot;
  *out = std::move(np);
}
}  // namespace

int pbh_texture_image(pbh_scene* s, int map_kind, const float map[12], const float* rgb, uint32_t w,
                      uint32_t h, int spectrum, int do_trilinear, float max_aniso, int wrap,
                      float scale, float gamma) {
  if (wrap < 0 || wrap > 2) {
    s->err = "image texture: bad wrap mode";
    return PBRTB200_EINVAL;
  }
  // texel conversion (imagemap.rs:108-121 / 163-176)
  std::vector<Tex3> px;
  size_t W = w, H = h;
  if (!rgb || w == 0 || h == 0) {
    const float v = std::pow(scale, gamma);
    px.push_back({v, v, v});
    W = H = 1;
  } else {
    px.resize(W * H);
    for (size_t i = 0; i < W * H; ++i) {
      const float r = rgb[3 * i], g = rgb[3 * i + 1], b = rgb[3 * i + 2];
      if (spectrum) {
        px[i] = {std::pow(r * scale, gamma), std::pow(g * scale, gamma), std::pow(b * scale, gamma)};
      } else {
        const float y = 0.212671f * r + 0.715160f * g + 0.072169f * b;  // Spectrum::y, spectrum.rs:37-41,468
        const float v = std::pow(y * scale, gamma);
        px[i] = {v, v, v};
      }
    }
  }
  // MIPMap::new (mipmap.rs:159-204)
  if ((W & (W - 1)) || (H & (H - 1))) {
    std::vector<Tex3> pot;
    mip_resize_pot(W, H, px, wrap, &W, &H, &pot);
    px.swap(pot);
  }
  pbrtb200_mipmap mm{};
  mm.width = (uint32_t)W;
  mm.height = (uint32_t)H;
  mm.do_trilinear = do_trilinear ? 1u : 0u;
  mm.max_anisotropy = max_aniso;
  mm.wrap = (uint32_t)wrap;
  mm.texel_offset = s->texels.size() / 4;
  auto append = [&](const std::vector<Tex3>& lv) {
    for (const Tex3& t : lv) {
      s->texels.push_back(t.r);
      s->texels.push_back(t.g);
      s->texels.push_back(t.b);
      s->texels.push_back(0.0f);
    }
  };
  append(px);
  uint32_t levels = 0;  // ulog2(max(w, h)) = bits - leading_zeros (mipmap.rs:108-110)
  for (size_t m = std::max(W, H); m; m >>= 1) ++levels;
  size_t lw = W, lh = H;
  std::vector<Tex3> last = std::move(px);
  for (uint32_t i = 1; i < levels; ++i) {
    const size_t nw = std::max<size_t>(lw / 2, 1), nh = std::max<size_t>(lh / 2, 1);
    std::vector<Tex3> nl(nw * nh);
    auto at = [&](int32_t si, int32_t ti) -> Tex3 {  // texel_at (mipmap.rs:112-138)
      const int32_t a = mip_wrap(si, (int32_t)lw, wrap, false), b = mip_wrap(ti, (int32_t)lh, wrap, false);
      if (a < 0 || b < 0) return Tex3{0.f, 0.f, 0.f};
      return last[(size_t)b * lw + (size_t)a];
    };
    for (int32_t t = 0; t < (int32_t)nh; ++t)
      for (int32_t si = 0; si < (int32_t)nw; ++si)
        nl[(size_t)t * nw + (size_t)si] =
            (((at(2 * si, 2 * t) + at(2 * si + 1, 2 * t)) + at(2 * si, 2 * t + 1)) + at(2 * si + 1, 2 * t + 1)) * 0.25f;
    append(nl);
    last.swap(nl);
    lw = nw;
    lh = nh;
  }
  mm.n_levels = levels;
  s->mipmaps.push_back(mm);
  pbrtb200_texture t{};
  t.kind = PBRTB200_TEX_IMAGE;
  t.map_kind = map_kind;
  std::memcpy(t.map, map, 48);
  t.tex1 = (int32_t)s->mipmaps.size() - 1;
  s->textures.push_back(t);
  return (int)s->textures.size() - 1;
}

int pbh_material_matte(pbh_scene* s, int kd, int sigma, int bump_map) {
  pbrtb200_material m{};
  m.bump = bump_map < 0 ? 0 : bump_map + 1;
  m.kind = PBRTB200_MAT_MATTE;
  m.kd = kd;
  m.sigma = sigma;
  s->materials.push_back(m);
  return (int)s->materials.size() - 1;
}
int pbh_material_plastic(pbh_scene* s, int kd, int ks, int roughness, int bump_map) {
  pbrtb200_material m{};
  m.bump = bump_map < 0 ? 0 : bump_map + 1;
  m.kind = PBRTB200_MAT_PLASTIC;
  m.kd = kd;
  m.ks = ks;
  m.roughness = roughness;
  s->materials.push_back(m);
  return (int)s->materials.size() - 1;
}
int pbh_light_point(pbh_scene* s, const float l2w[16], const float l2w_inv[16], const float I[3]) {
  pbrtb200_light l{};
  l.kind = PBRTB200_LIGHT_POINT;
  const Vec p = xf_point(l2w, {0.f, 0.f, 0.f});  // point.rs:22
  l.pos[0] = p.x;
  l.pos[1] = p.y;
  l.pos[2] = p.z;
  std::memcpy(l.intensity, I, 12);
  rows3(l2w_inv, l.w2l);
  l.num_samples = 1;
  s->lights.push_back(l);
  return (int)s->lights.size() - 1;
}
int pbh_light_spot(pbh_scene* s, const float l2w[16], const float l2w_inv[16], const float I[3],
                   float width_deg, float falloff_deg) {
  const int id = pbh_light_point(s, l2w, l2w_inv, I);
  pbrtb200_light& l = s->lights[(size_t)id];
  l.kind = PBRTB200_LIGHT_SPOT;
  l.cos_total_width = std::cos(radians(width_deg));      // spot.rs:31-32
  l.cos_falloff_start = std::cos(radians(falloff_deg));
  return id;
}
int pbh_light_area(pbh_scene* s, const float L[3], int num_samples) {
  pbrtb200_light l{};
  l.kind = PBRTB200_LIGHT_AREA;
  std::memcpy(l.intensity, L, 12);
  l.num_samples = num_samples;
  s->lights.push_back(l);
  return (int)s->lights.size() - 1;
}

int pbh_add_triangle_mesh(pbh_scene* s, const float o2w[16], const float o2w_inv[16], int ro,
                          const uint32_t* vi, uint64_t n_vi, const float* P, uint64_t n_p,
                          const float* N, const float* S, const float* UV, int material,
                          int area_light) {
  if (n_vi % 3 != 0) {  // mesh.rs:303 assert!(vi.len() % 3 == 0)
    s->err = "triangle_mesh: vi.len() % 3 != 0";
    return PBRTB200_EINVAL;
  }
  for (uint64_t i = 0; i < n_vi; ++i)
    if (vi[i] >= n_p) {
      s->err = "triangle_mesh: vertex index out of bounds";
      return PBRTB200_EINVAL;
    }
  MeshRec m;
  std::memcpy(m.o2w, o2w, 64);
  std::memcpy(m.o2w_inv, o2w_inv, 64);
  m.ro = ro != 0;
  m.vi.assign(vi, vi + n_vi);
  m.p.resize(n_p);
  for (uint64_t i = 0; i < n_p; ++i) m.p[i] = xf_point(o2w, {P[3 * i], P[3 * i + 1], P[3 * i + 2]});
  if (N) m.n.assign(N, N + 3 * n_p);
  if (S) m.s.assign(S, S + 3 * n_p);
  if (UV) m.uv.assign(UV, UV + 2 * n_p);
  m.material = material;
  m.area_light = area_light;
  s->objects.push_back({0, (uint32_t)s->meshes.size()});
  s->meshes.push_back(std::move(m));
  s->built = false;
  return (int)s->objects.size() - 1;
}

int pbh_add_sphere(pbh_scene* s, const float o2w[16], const float o2w_inv[16], int ro, float rad,
                   float z0, float z1, float pm, int material) {
  auto clampf = [](float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); };
  SphereRec r;
  std::memcpy(r.o2w, o2w, 64);
  std::memcpy(r.o2w_inv, o2w_inv, 64);
  r.ro = ro != 0;
  r.radius = rad;  // sphere.rs:28-44
  r.z_min = clampf(std::fmin(z0, z1), -rad, rad);
  r.z_max = clampf(std::fmax(z0, z1), -rad, rad);
  r.theta_min = std::acos(r.z_min / rad);
  r.theta_max = std::acos(r.z_max / rad);
  r.phi_max = radians(clampf(pm, 0.0f, 360.0f));
  r.material = material;
  s->objects.push_back({1, (uint32_t)s->spheres.size()});
  s->spheres.push_back(r);
  s->built = false;
  return (int)s->objects.size() - 1;
}

int pbh_add_cylinder(pbh_scene* s, const float o2w[16], const float o2w_inv[16], int ro, float rad,
                     float z0, float z1, float pm, int material) {
  auto clampf = [](float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); };
  SphereRec r;
  std::memcpy(r.o2w, o2w, 64);
  std::memcpy(r.o2w_inv, o2w_inv, 64);
  r.ro = ro != 0;
  r.kind = PBRTB200_QUADRIC_CYLINDER;  // cylinder.rs:27-38
  r.radius = rad;
  r.z_min = std::fmin(z0, z1);
  r.z_max = std::fmax(z0, z1);
  r.theta_min = r.theta_max = 0.0f;
  r.phi_max = radians(clampf(pm, 0.0f, 360.0f));
  r.material = material;
  s->objects.push_back({1, (uint32_t)s->spheres.size()});
  s->spheres.push_back(r);
  s->built = false;
  return (int)s->objects.size() - 1;
}

int pbh_add_disk(pbh_scene* s, const float o2w[16], const float o2w_inv[16], int ro, float height,
                 float radius, float inner_radius, float pm, int material) {
  auto clampf = [](float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); };
  SphereRec r;
  std::memcpy(r.o2w, o2w, 64);
  std::memcpy(r.o2w_inv, o2w_inv, 64);
  r.ro = ro != 0;
  r.kind = PBRTB200_QUADRIC_DISK;  // disk.rs:24-35
  r.radius = radius;
  r.z_min = r.z_max = height;  // object bound (-r, -r, h)..(r, r, h), disk.rs:77-81
  r.theta_min = inner_radius;
  r.theta_max = 0.0f;
  r.phi_max = radians(clampf(pm, 0.0f, 360.0f));
  r.material = material;
  s->objects.push_back({1, (uint32_t)s->spheres.size()});
  s->spheres.push_back(r);
  s->built = false;
  return (int)s->objects.size() - 1;
}

int pbh_build_bvh(pbh_scene* s, uint32_t max_prims, const char* split_method) {
  s->built = false;
  s->err.clear();
  int method = 2;
  if (split_method && std::strcmp(split_method, "middle") == 0)
    method = 0;
  else if (split_method && std::strcmp(split_method, "equal") == 0)
    method = 1;
  else if (!split_method || std::strcmp(split_method, "sah") != 0)
    std::printf("Warning: BVH split method %s unknown. Using \"SAH\"\n",
                split_method ? split_method : "(null)");  // bvh.rs:345-347

  // Refinement (primitive/mod.rs:48-62, mesh.rs:324-335): per object in input
  // order; a mesh contributes its triangles in ORIGINAL triple order with reversed vertex order.
  std::vector<PrimRef> refs;
  for (uint32_t o = 0; o < s->objects.size(); ++o) {
    const Object& ob = s->objects[o];
    if (ob.kind == 1) {
      refs.push_back({1, o, 0});
    } else {
      const size_t nt = s->meshes[ob.index].vi.size() / 3;
      for (size_t j = 0; j < nt; ++j) refs.push_back({0, o, (uint32_t)j});
    }
  }
  if (refs.empty()) {
    s->err = "scene has no primitives";
    return PBRTB200_EINVAL;
  }
  const size_t N = refs.size();
  std::vector<Box> bounds(N);
  std::vector<Vec> cent(N);
  for (size_t i = 0; i < N; ++i) {
    Box b;
    const Object& ob = s->objects[refs[i].object];
    if (refs[i].kind == 0) {
      const MeshRec& m = s->meshes[ob.index];
      const uint32_t j = refs[i].index;
      b.grow(m.p[m.vi[3 * j + 2]]);  // world_bound: BBox::new() U p1 U p2 U p3 (mesh.rs:195-204)
      b.grow(m.p[m.vi[3 * j + 1]]);
      b.grow(m.p[m.vi[3 * j]]);
    } else {
      b = sphere_world_bound(s->spheres[ob.index]);
    }
    bounds[i] = b;
    cent[i] = {(b.lo[0] + b.hi[0]) * 0.5f, (b.lo[1] + b.hi[1]) * 0.5f, (b.lo[2] + b.hi[2]) * 0.5f};
  }
  std::vector<uint32_t> idx(N);
  for (size_t i = 0; i < N; ++i) idx[i] = (uint32_t)i;
  s->nodes.clear();
  s->nodes.reserve(2 * N);
  Builder bld{bounds, cent, idx, {}, s->nodes, max_prims, method, &s->err};
  if (!bld.build(0, N) || !s->err.empty()) return PBRTB200_EINVAL;
  s->ordered.resize(N);
  for (size_t i = 0; i < N; ++i) s->ordered[i] = refs[idx[i]];

  // ---- flatten (the Rust shim's job) ----
  s->leaf_prim.clear();
  s->tris.clear();
  s->fspheres.clear();
  s->sphere_o2w.clear();
  s->fmeshes.clear();
  s->tri_uv.clear();
  s->tri_n.clear();
  s->tri_s.clear();
  s->area_prims.clear();
  bool any_uv = false, any_n = false, any_s = false;
  std::vector<uint32_t> mesh_of_object(s->objects.size(), 0);
  for (uint32_t o = 0; o < s->objects.size(); ++o) {
    if (s->objects[o].kind != 0) continue;
    const MeshRec& m = s->meshes[s->objects[o].index];
    pbrtb200_mesh fm{};
    rows3(m.o2w, fm.o2w);
    rows3(m.o2w_inv, fm.o2w_inv);
    fm.material = (uint32_t)std::max(m.material, 0);
    fm.area_light = m.area_light;
    fm.flip = (m.ro ^ swaps_handedness(m.o2w)) ? 1u : 0u;
    fm.has_uv = !m.uv.empty();
    fm.has_n = !m.n.empty();
    fm.has_s = !m.s.empty();
    any_uv |= fm.has_uv != 0;
    any_n |= fm.has_n != 0;
    any_s |= fm.has_s != 0;
    mesh_of_object[o] = (uint32_t)s->fmeshes.size();
    s->fmeshes.push_back(fm);
  }
  const bool has_spheres = !s->spheres.empty();
  std::vector<uint32_t> sphere_slot(s->objects.size(), 0);
  for (uint32_t o = 0; o < s->objects.size(); ++o) {
    if (s->objects[o].kind != 1) continue;
    const SphereRec& r = s->spheres[s->objects[o].index];
    pbrtb200_sphere80 f{};
    rows3(r.o2w_inv, f.w2o);
    f.radius = r.radius;
    f.z_min = r.z_min;
    f.z_max = r.z_max;
    f.phi_max = r.phi_max;
    f.theta_min = r.theta_min;
    f.theta_max = r.theta_max;
    f.material = (uint32_t)std::max(r.material, 0);
    f.flip = ((r.ro ^ swaps_handedness(r.o2w)) ? 1u : 0u) | (r.kind << PBRTB200_QUADRIC_KIND_SHIFT);
    sphere_slot[o] = (uint32_t)s->fspheres.size();
    s->fspheres.push_back(f);
    float rows[12];
    rows3(r.o2w, rows);
    s->sphere_o2w.insert(s->sphere_o2w.end(), rows, rows + 12);
  }
  s->tris.reserve(N);
  if (has_spheres) s->leaf_prim.reserve(N);
  for (size_t i = 0; i < N; ++i) {
    const PrimRef& pr = s->ordered[i];
    if (pr.kind == 1) {
      s->leaf_prim.push_back(0x80000000u | sphere_slot[pr.object]);
      continue;
    }
    const MeshRec& m = s->meshes[s->objects[pr.object].index];
    const uint32_t j = pr.index;
    const uint32_t v[3] = {m.vi[3 * j + 2], m.vi[3 * j + 1], m.vi[3 * j]};  // mesh.rs:329-331
    pbrtb200_tri48 t{};
    const Vec p1 = m.p[v[0]], p2 = m.p[v[1]], p3 = m.p[v[2]];
    t.p1[0] = p1.x; t.p1[1] = p1.y; t.p1[2] = p1.z;
    t.p2[0] = p2.x; t.p2[1] = p2.y; t.p2[2] = p2.z;
    t.p3[0] = p3.x; t.p3[1] = p3.y; t.p3[2] = p3.z;
    t.mesh = mesh_of_object[pr.object];
    t.attr = (uint32_t)s->tris.size();
    t.user = j;
    if (has_spheres) s->leaf_prim.push_back((uint32_t)s->tris.size());
    if (any_uv)
      for (int k = 0; k < 3; ++k) {
        s->tri_uv.push_back(m.uv.empty() ? 0.f : m.uv[2 * v[k]]);
        s->tri_uv.push_back(m.uv.empty() ? 0.f : m.uv[2 * v[k] + 1]);
      }
    if (any_n)
      for (int k = 0; k < 3; ++k)
        for (int c = 0; c < 3; ++c) s->tri_n.push_back(m.n.empty() ? 0.f : m.n[3 * v[k] + c]);
    if (any_s)
      for (int k = 0; k < 3; ++k)
        for (int c = 0; c < 3; ++c) s->tri_s.push_back(m.s.empty() ? 0.f : m.s[3 * v[k] + c]);
    s->tris.push_back(t);
  }
  // emissive triangles per area light, in refined (BVH-input) order
  {
    std::vector<uint32_t> pos_of_ref(N);
    for (size_t i = 0; i < N; ++i) pos_of_ref[idx[i]] = (uint32_t)i;
    for (auto& l : s->lights) {
      l.first_tri = 0;
      l.n_tris = 0;
    }
    for (size_t li = 0; li < s->lights.size(); ++li) {
      if (s->lights[li].kind != PBRTB200_LIGHT_AREA) continue;
      s->lights[li].first_tri = (uint32_t)s->area_prims.size();
      for (size_t r = 0; r < N; ++r) {
        if (refs[r].kind != 0) continue;
        const MeshRec& m = s->meshes[s->objects[refs[r].object].index];
        if (m.area_light == (int)li) s->area_prims.push_back(pos_of_ref[r]);
      }
      s->lights[li].n_tris = (uint32_t)s->area_prims.size() - s->lights[li].first_tri;
      if (s->lights[li].n_tris == 0) {
        s->err = "area light without emissive triangles";
        return PBRTB200_EINVAL;
      }
    }
  }
  pbrtb200_scene& f = s->flat;
  std::memset(&f, 0, sizeof f);
  f.nodes = s->nodes.data();
  f.n_nodes = (uint32_t)s->nodes.size();
  f.leaf_prim = has_spheres ? s->leaf_prim.data() : nullptr;
  f.n_prims = (uint32_t)N;
  f.tris = s->tris.data();
  f.n_tris = (uint32_t)s->tris.size();
  f.spheres = s->fspheres.data();
  f.sphere_o2w = s->sphere_o2w.data();
  f.n_spheres = (uint32_t)s->fspheres.size();
  f.meshes = s->fmeshes.data();
  f.n_meshes = (uint32_t)s->fmeshes.size();
  f.tri_uv = any_uv ? s->tri_uv.data() : nullptr;
  f.tri_n = any_n ? s->tri_n.data() : nullptr;
  f.tri_s = any_s ? s->tri_s.data() : nullptr;
  f.n_attr = (uint32_t)s->tris.size();
  f.materials = s->materials.data();
  f.n_materials = (uint32_t)s->materials.size();
  f.textures = s->textures.data();
  f.n_textures = (uint32_t)s->textures.size();
  f.lights = s->lights.data();
  f.n_lights = (uint32_t)s->lights.size();
  f.area_prims = s->area_prims.data();
  f.n_area_prims = (uint32_t)s->area_prims.size();
  f.mipmaps = s->mipmaps.data();
  f.n_mipmaps = (uint32_t)s->mipmaps.size();
  f.texels = s->texels.data();
  f.n_texels = s->texels.size() / 4;
  s->built = true;
  return PBRTB200_OK;
}

const pbrtb200_scene* pbh_flat_scene(const pbh_scene* s) { return s->built ? &s->flat : nullptr; }

void pbh_prim_order(const pbh_scene* s, uint32_t* out3) {
  for (size_t i = 0; i < s->ordered.size(); ++i) {
    out3[3 * i] = s->ordered[i].kind;
    out3[3 * i + 1] = s->ordered[i].object;
    out3[3 * i + 2] = s->ordered[i].index;
  }
}

int pbh_camera_perspective(const float cam2world[16], const float sw[4], float sopen, float sclose,
                           float lensr, float focald, float fov, int x_res, int y_res,
                           pbrtb200_camera* out) {
  // camera/mod.rs:105-135
  const float znear = 1e-2f, zfar = 1000.0f;
  Mat p = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, zfar / (zfar - znear), -(zfar * znear) / (zfar - znear), 0, 0, 1, 0};
  Mat p_inv;
  if (!mat_invert(p, p_inv)) return PBRTB200_ESINGULAR;
  const float inv_tan = 1.0f / std::tan(radians(fov) / 2.0f);
  Mat sc, sc_inv, persp, persp_inv;
  pbh_scale(inv_tan, inv_tan, 1.0f, sc, sc_inv);
  pbh_mul(sc, sc_inv, p, p_inv, persp, persp_inv);
  // projective.rs:55-63
  Mat a, ai, b, bi, c, ci, ab, abi, s2r, s2r_inv;
  pbh_scale((float)x_res, (float)y_res, 1.0f, a, ai);
  pbh_scale(1.0f / (sw[1] - sw[0]), 1.0f / (sw[2] - sw[3]), 1.0f, b, bi);
  const float tv[3] = {-sw[0], -sw[3], 0.0f};
  pbh_translate(tv, c, ci);
  pbh_mul(a, ai, b, bi, ab, abi);
  pbh_mul(ab, abi, c, ci, s2r, s2r_inv);
  // raster_to_camera = camera_to_screen.inverse() * raster_to_screen
  Mat r2c, r2c_inv;
  pbh_mul(persp_inv, persp, s2r_inv, s2r, r2c, r2c_inv);
  std::memcpy(out->raster_to_camera, r2c, 64);
  std::memcpy(out->camera_to_world, cam2world, 64);
  const Vec dx = xf_vector(r2c, {1, 0, 0}) - xf_vector(r2c, {0, 0, 0});  // mod.rs:129-133
  const Vec dy = xf_vector(r2c, {0, 1, 0}) - xf_vector(r2c, {0, 0, 0});
  out->dx_camera[0] = dx.x;
  out->dx_camera[1] = dx.y;
  out->dx_camera[2] = dx.z;
  out->dy_camera[0] = dy.x;
  out->dy_camera[1] = dy.y;
  out->dy_camera[2] = dy.z;
  out->shutter_open = sopen;
  out->shutter_close = sclose;
  out->lens_radius = lensr;
  out->focal_distance = focald;
  return PBRTB200_OK;
}

int pbh_film_image(int x_res, int y_res, int filter_type, float xw, float yw, float p0, float p1,
                   const float crop[4], pbrtb200_film* out) {
  // film.rs:77-122
  out->x_res = x_res;
  out->y_res = y_res;
  out->x_pixel_start = pbh::sat_i32(std::ceil((float)x_res * crop[0]));
  out->x_pixel_count = std::max(pbh::sat_i32(std::ceil((float)x_res * crop[1])) - out->x_pixel_start, 1);
  out->y_pixel_start = pbh::sat_i32(std::ceil((float)y_res * crop[2]));
  out->y_pixel_count = std::max(pbh::sat_i32(std::ceil((float)y_res * crop[3])) - out->y_pixel_start, 1);
  out->filter_xw = xw;
  out->filter_yw = yw;
  for (int y = 0; y < 16; ++y) {
    const float fy = ((float)y + 0.5f) * yw / 16.0f;
    for (int x = 0; x < 16; ++x) {
      const float fx = ((float)x + 0.5f) * xw / 16.0f;
      out->filter_table[y * 16 + x] = filter_eval(filter_type, xw, yw, p0, p1, fx, fy);
    }
  }
  return PBRTB200_OK;
}

void pbh_film_sample_extent(const pbrtb200_film* f, int32_t out[4]) {  // film.rs:271-289
  out[0] = pbh::sat_i32(std::floor((float)f->x_pixel_start + 0.5f - f->filter_xw));
  out[1] = pbh::sat_i32(std::floor((float)f->x_pixel_start + 0.5f + (float)f->x_pixel_count + f->filter_xw));
  out[2] = pbh::sat_i32(std::floor((float)f->y_pixel_start + 0.5f - f->filter_yw));
  out[3] = pbh::sat_i32(std::floor((float)f->y_pixel_start + 0.5f + (float)f->y_pixel_count + f->filter_yw));
}

uint32_t pbh_num_tasks(uint32_t num_cpus, uint32_t num_pixels) {  // sampler_renderer.rs:39-44
  const uint32_t x = std::max(32u * num_cpus, num_pixels / 256u);
  const uint32_t lz = (uint32_t)__builtin_clz(x);
  return 31u - lz + ((x & (x - 1u)) == 0u ? 0u : 1u);
}

void pbh_film_to_rgb(const float* xyzw, uint64_t n, float* rgb) {
  for (uint64_t i = 0; i < n; ++i) {
    const float* p = xyzw + 4 * i;
    float r = 3.240479f * p[0] - 1.37150f * p[1] - 0.498535f * p[2];  // spectrum.rs:31-35
    float g = -0.969256f * p[0] + 1.875991f * p[1] + 0.041556f * p[2];
    float b = 0.055648f * p[0] - 0.204043f * p[1] + 1.057311f * p[2];
    if (p[3] != 0.0f) {
      const float inv = 1.0f / p[3];
      r = std::fmax(r * inv, 0.0f);
      g = std::fmax(g * inv, 0.0f);
      b = std::fmax(b * inv, 0.0f);
    }
    rgb[3 * i] = r;
    rgb[3 * i + 1] = g;
    rgb[3 * i + 2] = b;
  }
}


void pbh_rgb_to_bytes(const float* rgb, uint64_t n, uint8_t* out) {  // film.rs:21-23
  for (uint64_t i = 0; i < n; ++i) {
    float v = 255.0f * std::pow(rgb[i], 1.0f / 2.2f) + 0.5f;
    v = v < 0.0f ? 0.0f : (v > 255.0f ? 255.0f : v);
    out[i] = std::isnan(v) ? (uint8_t)0 : (uint8_t)v;  // Rust `as u8`: saturating, NaN -> 0
  }
}

// 8-bit RGB PNG (what `image::ImageBuffer<Rgb<u8>>::save` writes for a .png name, film.rs:15-33):
// one IDAT, filter type 0 on every scanline, zlib deflate.
int pbh_write_png(const char* path, const uint8_t* rgb8, uint32_t width, uint32_t height) {
  if (!path || !rgb8 || width == 0 || height == 0) return PBRTB200_EINVAL;
  const size_t stride = (size_t)width * 3;
  std::vector<uint8_t> raw((stride + 1) * (size_t)height);
  for (uint32_t y = 0; y < height; ++y) {
    raw[(stride + 1) * y] = 0;
    std::memcpy(&raw[(stride + 1) * y + 1], rgb8 + stride * y, stride);
  }
  uLongf zlen = compressBound((uLong)raw.size());
  std::vector<uint8_t> z(zlen);
  if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return PBRTB200_ENOMEM;
  FILE* f = std::fopen(path, "wb");
  if (!f) return PBRTB200_EINVAL;
  auto be32 = [](uint8_t* p, uint32_t v) {
    p[0] = (uint8_t)(v >> 24);
    p[1] = (uint8_t)(v >> 16);
    p[2] = (uint8_t)(v >> 8);
    p[3] = (uint8_t)v;
  };
  auto chunk = [&](const char type[4], const uint8_t* data, uint32_t len) {
    uint8_t hdr[8];
    be32(hdr, len);
    std::memcpy(hdr + 4, type, 4);
    std::fwrite(hdr, 1, 8, f);
    if (len) std::fwrite(data, 1, len, f);
    uLong c = crc32(0L, hdr + 4, 4);
    if (len) c = crc32(c, data, len);
    uint8_t crc[4];
    be32(crc, (uint32_t)c);
    std::fwrite(crc, 1, 4, f);
  };
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  std::fwrite(sig, 1, 8, f);
  uint8_t ihdr[13];
  be32(ihdr, width);
  be32(ihdr + 4, height);
  ihdr[8] = 8;   // bit depth
  ihdr[9] = 2;   // colour type: truecolour
  ihdr[10] = ihdr[11] = ihdr[12] = 0;
  chunk("IHDR", ihdr, 13);
  chunk("IDAT", z.data(), (uint32_t)zlen);
  chunk("IEND", nullptr, 0);
  const bool ok = std::ferror(f) == 0;
  return (std::fclose(f) == 0 && ok) ? PBRTB200_OK : PBRTB200_EINVAL;
}

// Linear float RGB as a little-endian PFM ("PF", bottom-to-top scanlines): the lossless companion
// of the 8-bit file (the crate lists `exr` in Cargo.toml but never calls it).
int pbh_write_pfm(const char* path, const float* rgb, uint32_t width, uint32_t height) {
  if (!path || !rgb || width == 0 || height == 0) return PBRTB200_EINVAL;
  FILE* f = std::fopen(path, "wb");
  if (!f) return PBRTB200_EINVAL;
  std::fprintf(f, "PF\n%u %u\n-1.0\n", width, height);
  for (uint32_t y = height; y-- > 0;) std::fwrite(rgb + (size_t)y * width * 3, sizeof(float), (size_t)width * 3, f);
  const bool ok = std::ferror(f) == 0;
  return (std::fclose(f) == 0 && ok) ? PBRTB200_OK : PBRTB200_EINVAL;
}
}  // extern "C"
