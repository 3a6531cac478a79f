// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// Shapes (Triangle, Sphere), primitive refinement order, the BVH builder and the closest-hit
// traversal of pbrt_rust, restated line-faithfully.  Citations are relative to the reference root.
#pragma once
#include <algorithm>
#include <memory>

#include "geom.hpp"

namespace orc {

// shape/mod.rs:33-54 (ShapeBase; the global shape_id counter is irrelevant to results)
struct ShapeBase {
  Transform o2w, w2o;
  bool reverse_orientation = false;
  bool swaps_handedness = false;
  ShapeBase() = default;
  ShapeBase(const Transform& a, const Transform& b, bool ro)
      : o2w(a), w2o(b), reverse_orientation(ro), swaps_handedness(a.swaps_handedness()) {}
};

// shape/mesh.rs:276-321.  Mesh::new pre-transforms P to world space with o2w (":308").
struct Mesh {
  ShapeBase base;
  std::vector<uint32_t> vi;
  std::vector<V3> p;   // world space
  std::vector<V3> n;   // optional (object space), empty if absent
  std::vector<V3> s;   // optional
  std::vector<float> uvs;  // optional, 2 per vertex
  uint32_t material = 0;
  int32_t area_light = -1;  // index into Scene::area_lights, -1 = not emissive (oracle-defined, D9)
  Mesh(const Transform& o2w, const Transform& w2o, bool ro, const uint32_t* idx, size_t n_idx,
       const float* P, size_t n_p, const float* N, const float* S, const float* UV)
      : base(o2w, w2o, ro) {
    if (n_idx % 3 != 0) throw std::runtime_error("vi.len() % 3 != 0");
    vi.assign(idx, idx + n_idx);
    p.resize(n_p);
    for (size_t i = 0; i < n_p; ++i) p[i] = o2w.pt(V3(P[3 * i], P[3 * i + 1], P[3 * i + 2]));
    if (N) {
      n.resize(n_p);
      for (size_t i = 0; i < n_p; ++i) n[i] = V3(N[3 * i], N[3 * i + 1], N[3 * i + 2]);
    }
    if (S) {
      s.resize(n_p);
      for (size_t i = 0; i < n_p; ++i) s[i] = V3(S[3 * i], S[3 * i + 1], S[3 * i + 2]);
    }
    if (UV) uvs.assign(UV, UV + 2 * n_p);
  }
};

// shape/sphere.rs:16-44
// (Also carries the crate's other two quadrics, so that the primitive plumbing stays one type:
// shape 1 = Cylinder (shape/cylinder.rs:17-38), shape 2 = Disk (shape/disk.rs:15-35).)
struct Sphere {
  ShapeBase base;
  float radius, phi_max, z_min, z_max, theta_min, theta_max;
  uint32_t material = 0;
  int shape = 0;                           // 0 sphere, 1 cylinder, 2 disk
  float height = 0.f, inner_radius = 0.f;  // disk
  struct CylinderTag {};
  struct DiskTag {};
  // Cylinder::new (cylinder.rs:27-38)
  Sphere(CylinderTag, const Transform& o2w, const Transform& w2o, bool ro, float rad, float z0, float z1, float pm)
      : base(o2w, w2o, ro) {
    shape = 1;
    radius = rad;
    z_min = rmin(z0, z1);
    z_max = rmax(z0, z1);
    theta_min = theta_max = 0.f;
    phi_max = as_radians(rclamp(pm, 0.0f, 360.0f));
  }
  // Disk::new (disk.rs:24-35)
  Sphere(DiskTag, const Transform& o2w, const Transform& w2o, bool ro, float ht, float r, float ri, float tmax)
      : base(o2w, w2o, ro) {
    shape = 2;
    height = ht;
    radius = r;
    inner_radius = ri;
    z_min = z_max = ht;
    theta_min = theta_max = 0.f;
    phi_max = as_radians(rclamp(tmax, 0.0f, 360.0f));
  }
  // cylinder.rs:110-113 / disk.rs:83-87 / sphere.rs:119-121
  float area() const {
    if (shape == 1) return (z_max - z_min) * phi_max * radius;
    if (shape == 2) return 0.5f * phi_max * (radius * radius - inner_radius * inner_radius);
    return phi_max * radius * (z_max - z_min);
  }
  Sphere(const Transform& o2w, const Transform& w2o, bool ro, float rad, float z0, float z1,
         float pm)
      : base(o2w, w2o, ro) {
    float zmin = rclamp(rmin(z0, z1), -rad, rad);
    float zmax = rclamp(rmax(z0, z1), -rad, rad);
    radius = rad;
    z_min = zmin;
    z_max = zmax;
    theta_min = std::acos(zmin / rad);
    theta_max = std::acos(zmax / rad);
    phi_max = as_radians(rclamp(pm, 0.0f, 360.0f));
  }
  // sphere.rs:112-117
  BBox object_bound() const {  // also cylinder.rs:104-108; disk.rs:77-81 (z = height on both corners)
    return BBox(V3(-radius, -radius, z_min), V3(radius, radius, z_max));
  }
  // sphere.rs:124-128
  BBox world_bound() const { return xf_bbox(base.o2w, object_bound()); }

  // sphere.rs:46-107; `ray` already in object space.
  bool intersection_point(const Ray& ray, float* t_out, float* phi_out) const {
    if (shape == 1) return cylinder_point(ray, t_out, phi_out);
    if (shape == 2) return disk_point(ray, t_out, phi_out);
    float a = length_squared(ray.d);
    float b = 2.0f * dot(ray.d, ray.o);
    float c = length_squared(ray.o) - radius * radius;
    float t0, t1;
    if (!quadratic(a, b, c, &t0, &t1)) return false;
    if (t0 > ray.maxt || t1 < ray.mint) return false;
    float t_hit = t0;
    if (t0 < ray.mint) {
      t_hit = t1;
      if (t_hit > ray.maxt) return false;
    }
    auto get_hit = [&](float t, V3* hit, float* angle) {
      V3 h = ray.at(t);
      if (h.x == 0.0f && h.y == 0.0f) h.x = 1e-5f * radius;
      float ang = std::atan2(h.y, h.x);
      if (ang < 0.0f) ang += 2.0f * PI_F;
      *hit = h;
      *angle = ang;
    };
    auto invalid = [&](const V3& h, float ang) {
      return (h.z > -radius && h.z < z_min) || (h.z < radius && h.z > z_max) || (ang > phi_max);
    };
    V3 h;
    float ang;
    get_hit(t_hit, &h, &ang);
    if (invalid(h, ang)) {
      if (t_hit == t1) return false;
      if (t1 > ray.maxt) return false;
      t_hit = t1;
      get_hit(t_hit, &h, &ang);
      if (invalid(h, ang)) return false;
    }
    *t_out = t_hit;
    *phi_out = ang;
    return true;
  }
  // cylinder.rs:40-100
  bool cylinder_point(const Ray& r, float* t_out, float* phi_out) const {
    float a = r.d.x * r.d.x + r.d.y * r.d.y;
    float b = 2.0f * (r.d.x * r.o.x + r.d.y * r.o.y);
    float c = r.o.x * r.o.x + r.o.y * r.o.y - radius * radius;
    float t0, t1;
    if (!quadratic(a, b, c, &t0, &t1)) return false;
    if (t0 > r.maxt || t1 < r.mint) return false;
    float t_hit = t0;
    if (t0 < r.mint) {
      t_hit = t1;
      if (t_hit > r.maxt) return false;
    }
    auto get_hit = [&](float t, V3* hit, float* angle) {
      V3 h = r.at(t);
      if (h.x == 0.0f && h.y == 0.0f) h.x = 1e-5f * radius;
      float ang = std::atan2(h.y, h.x);
      if (ang < 0.0f) ang = ang + 2.0f * PI_F;
      *hit = h;
      *angle = ang;
    };
    auto invalid = [&](const V3& h, float ang) { return h.z < z_min || h.z > z_max || ang > phi_max; };
    V3 h;
    float ang;
    get_hit(t_hit, &h, &ang);
    if (invalid(h, ang)) {
      if (t_hit == t1) return false;
      if (t1 > r.maxt) return false;
      t_hit = t1;
      get_hit(t_hit, &h, &ang);
      if (invalid(h, ang)) return false;
    }
    *t_out = t_hit;
    *phi_out = ang;
    return true;
  }
  // disk.rs:37-71
  bool disk_point(const Ray& r, float* t_out, float* phi_out) const {
    if (std::fabs(r.d.z) < 1e-6f) return false;
    float t_hit = (height - r.o.z) / r.d.z;
    if (t_hit < r.mint || t_hit > r.maxt) return false;
    V3 p_hit = r.at(t_hit);
    float dist2 = p_hit.x * p_hit.x + p_hit.y * p_hit.y;
    if (dist2 > (radius * radius) || dist2 < (inner_radius * inner_radius)) return false;
    float a = std::atan2(p_hit.y, p_hit.x);
    float phi = a < 0.0f ? a + 2.0f * PI_F : a;
    if (phi > phi_max) return false;
    *t_out = t_hit;
    *phi_out = phi;
    return true;
  }
};

// diff_geom.rs:12-79
struct DiffGeom {
  V3 p, nn;
  float u = 0.f, v = 0.f;
  V3 dpdu, dpdv, dndu, dndv;
  V3 dpdx, dpdy;
  float dudx = 0.f, dudy = 0.f, dvdx = 0.f, dvdy = 0.f;
  bool flip = false;  // shape.reverse_orientation ^ shape.transform_swaps_handedness
  DiffGeom() = default;
  // diff_geom.rs:52-79
  DiffGeom(const V3& p_, const V3& dpdu_, const V3& dpdv_, const V3& dndu_, const V3& dndv_,
           float u_, float v_, const ShapeBase* shape)
      : p(p_), u(u_), v(v_), dpdu(dpdu_), dpdv(dpdv_), dndu(dndu_), dndv(dndv_) {
    V3 norm = normalize(cross(dpdu_, dpdv_));
    if (shape) {
      flip = shape->reverse_orientation ^ shape->swaps_handedness;
      if (flip) norm = norm * -1.f;
    }
    nn = norm;
  }
  // diff_geom.rs:81-152
  void compute_differentials(const RayDifferential& ray) {
    if (!ray.has_differentials) {
      dpdx = dpdy = V3();
      dudx = dudy = dvdx = dvdy = 0.f;
      return;
    }
    V3 nvec = nn;
    float d = -(dot(nvec, p));
    V3 px, py;
    {
      float ndrx = -(dot(nvec, ray.rx_origin) + d);
      float ndrd = dot(nvec, ray.rx_dir);
      float tx = ndrx / ndrd;
      px = ray.rx_origin + tx * ray.rx_dir;
    }
    {
      float ndry = -(dot(nvec, ray.ry_origin) + d);
      float ndrd = dot(nvec, ray.ry_dir);
      float ty = ndry / ndrd;
      py = ray.ry_origin + ty * ray.ry_dir;
    }
    dpdx = px - p;
    dpdy = py - p;
    int ax0, ax1;
    if (std::fabs(nn.x) > std::fabs(nn.y) && std::fabs(nn.x) > std::fabs(nn.z)) {
      ax0 = 1;
      ax1 = 2;
    } else if (std::fabs(nn.y) > std::fabs(nn.z)) {
      ax0 = 0;
      ax1 = 2;
    } else {
      ax0 = 0;
      ax1 = 1;
    }
    float a[2][2] = {{dpdu[ax0], dpdv[ax0]}, {dpdu[ax1], dpdv[ax1]}};
    float bx[2] = {dpdx[ax0], dpdx[ax1]};
    float by[2] = {dpdy[ax0], dpdy[ax1]};
    if (!solve_linear_system_2x2(a, bx, &dudx, &dvdx)) dudx = dvdx = 0.f;
    if (!solve_linear_system_2x2(a, by, &dudy, &dvdy)) dudy = dvdy = 0.f;
  }
};

// One fully refined primitive (primitive/mod.rs:27-231 + primitive/geometric.rs): either a
// Triangle (shape/mesh.rs:27-30) or a Sphere.
struct Prim {
  enum Kind : uint8_t { TRI = 0, SPH = 1 } kind = TRI;
  const Mesh* mesh = nullptr;
  uint32_t v[3] = {0, 0, 0};  // vertex indices (already in the refine-reversed order)
  const Sphere* sphere = nullptr;
  uint32_t source_index = 0;  // triple index j within its mesh, or sphere ordinal
  uint32_t source_object = 0; // ordinal of the user-level primitive it came from
  // mesh.rs:195-204 / sphere.rs:124-128
  BBox world_bound() const {
    if (kind == TRI)
      return BBox().united(mesh->p[v[0]]).united(mesh->p[v[1]]).united(mesh->p[v[2]]);
    return sphere->world_bound();
  }
  uint32_t material() const { return kind == TRI ? mesh->material : sphere->material; }
};

// mesh.rs:41-72  Möller–Trumbore exactly as written.
inline bool tri_intersection_point(const V3& p1, const V3& p2, const V3& p3, const Ray& r,
                                   float* t_out, float* b1_out, float* b2_out) {
  V3 e1 = p2 - p1;
  V3 e2 = p3 - p1;
  V3 s1 = cross(r.d, e2);
  float divisor = dot(s1, e1);
  if (divisor == 0.f) return false;
  float inv_divisor = 1.0f / divisor;
  V3 s = r.o - p1;
  float b1 = dot(s1, s) * inv_divisor;
  if (b1 < 0.0f || b1 > 1.0f) return false;
  V3 s2 = cross(s, e1);
  float b2 = dot(r.d, s2) * inv_divisor;
  if (b2 < 0.0f || (b1 + b2) > 1.0f) return false;
  float t = dot(e2, s2) * inv_divisor;
  if (t < r.mint || t > r.maxt) return false;
  *t_out = t;
  *b1_out = b1;
  *b2_out = b2;
  return true;
}

struct Hit {
  uint32_t prim = 0xFFFFFFFFu;  // index into the BVH's ordered primitive list (leaf order)
  float t = 0.f;
  float b1 = 0.f, b2 = 0.f;  // triangle barycentrics; for spheres b1 = phi
};

struct TraceCounters {
  uint64_t nodes_visited = 0;  // nodes popped (incl. those failing the box test), bvh.rs:390-395
  uint64_t tris_tested = 0;
  uint64_t spheres_tested = 0;
};

// Shape::intersect restricted to (t, params): geometric.rs:62-73 sets ray.maxt = t_hit on a hit.
inline bool prim_intersect(const Prim& pr, const Ray& ray, Hit* h, TraceCounters* tc) {
  if (pr.kind == Prim::TRI) {
    if (tc) tc->tris_tested++;
    float t, b1, b2;
    const Mesh& m = *pr.mesh;
    if (!tri_intersection_point(m.p[pr.v[0]], m.p[pr.v[1]], m.p[pr.v[2]], ray, &t, &b1, &b2))
      return false;
    ray.maxt = t;
    h->t = t;
    h->b1 = b1;
    h->b2 = b2;
    return true;
  }
  if (tc) tc->spheres_tested++;
  Ray o = xf_ray(pr.sphere->base.w2o, ray);  // sphere.rs:137
  float t, phi;
  if (!pr.sphere->intersection_point(o, &t, &phi)) return false;
  ray.maxt = t;
  h->t = t;
  h->b1 = phi;
  h->b2 = 0.f;
  return true;
}

// ---------------------------------------------------------------------------------------------
// primitive/aggregates/bvh.rs
enum class SplitMethod { Middle, EqualCounts, SAH };

struct PackedNode {  // bvh.rs:260-272
  BBox bounds;
  bool leaf = false;
  uint32_t offset = 0;  // Leaf: prim_offset ; Inner: second_child_offset
  uint32_t count = 0;   // Leaf: num_prims    ; Inner: axis
};

struct BVH {
  std::vector<PackedNode> nodes;
  std::vector<Prim> prims;  // ordered primitives (bvh.rs:331)

  struct Info {  // bvh.rs:66-84
    uint32_t prim;
    V3 centroid;
    BBox bounds;
  };
  struct BuildNode {  // bvh.rs:21-36
    BBox bounds;
    bool leaf;
    uint32_t first = 0, n = 0;
    std::unique_ptr<BuildNode> c1, c2;
    int axis = 0;
    uint32_t num_nodes = 1;
  };
  static constexpr int NUM_BUCKETS = 12;  // bvh.rs:188

  const std::vector<Prim>* src = nullptr;
  size_t max_prims = 1;
  SplitMethod sm = SplitMethod::SAH;

  // bvh.rs:334-362
  void build(const std::vector<Prim>& input, size_t mp, SplitMethod method) {
    src = &input;
    max_prims = mp;
    sm = method;
    nodes.clear();
    prims.clear();
    if (input.empty()) return;  // the reference would recurse on an empty vec; callers never do
    std::vector<Info> data(input.size());
    for (size_t i = 0; i < input.size(); ++i) {
      BBox b = input[i].world_bound();
      data[i].prim = (uint32_t)i;
      data[i].bounds = b;
      data[i].centroid = (b.p_min + b.p_max) * 0.5f;  // bvh.rs:77
    }
    std::vector<uint32_t> ordered;
    ordered.reserve(input.size());
    std::unique_ptr<BuildNode> root = recursive_build(std::move(data), ordered);
    nodes.reserve(root->num_nodes);
    flatten(*root);
    prims.reserve(ordered.size());
    for (uint32_t i : ordered) prims.push_back(input[i]);
  }

  static std::unique_ptr<BuildNode> make_leaf(const BBox& b, const std::vector<Info>& v,
                                              std::vector<uint32_t>& ordered) {
    auto nd = std::make_unique<BuildNode>();
    nd->bounds = b;
    nd->leaf = true;
    nd->first = (uint32_t)ordered.size();  // == the reference's later `offset()` fix-up
    nd->n = (uint32_t)v.size();
    for (const Info& i : v) ordered.push_back(i.prim);
    return nd;
  }

  // bvh.rs:189-258
  std::unique_ptr<BuildNode> recursive_build(std::vector<Info> v, std::vector<uint32_t>& ordered) {
    BBox bbox;
    for (const Info& i : v) bbox = bbox.united(i.bounds);
    size_t num_prims = v.size();
    if (num_prims == 1) return make_leaf(bbox, v, ordered);
    BBox cb;
    for (const Info& i : v) cb = cb.united(i.centroid);
    int dim = cb.max_extent();
    if (cb.p_min[dim] == cb.p_max[dim]) return make_leaf(bbox, v, ordered);

    std::vector<Info> p1, p2;
    auto equal_counts = [&](std::vector<Info>& w) {  // bvh.rs:92-104
      partition_by(
          0, w.size(), [&](size_t i) { return w[i].centroid[dim]; },
          [&](size_t i, size_t j) { std::swap(w[i], w[j]); });
      size_t n = w.size();
      p1.assign(w.begin(), w.begin() + n / 2);
      p2.assign(w.begin() + n / 2, w.end());
    };
    switch (sm) {
      case SplitMethod::Middle: {  // bvh.rs:86-90
        float p_mid = 0.5f * (cb.p_min[dim] + cb.p_max[dim]);
        for (const Info& i : v) (i.centroid[dim] < p_mid ? p1 : p2).push_back(i);
        break;
      }
      case SplitMethod::EqualCounts:
        equal_counts(v);
        break;
      case SplitMethod::SAH: {  // bvh.rs:106-186
        if (num_prims <= 4) {
          equal_counts(v);
          break;
        }
        struct Bucket {
          size_t count = 0;
          BBox b;
        } buckets[NUM_BUCKETS];
        auto bucket_for = [&](const Info& p) -> size_t {
          float pdist = p.centroid[dim] - cb.p_min[dim];
          float dist = cb.p_max[dim] - cb.p_min[dim];
          size_t b = (size_t)f2usize((float)NUM_BUCKETS * (pdist / dist));
          return b == (size_t)NUM_BUCKETS ? (size_t)NUM_BUCKETS - 1 : b;
        };
        for (const Info& p : v) {
          size_t b = bucket_for(p);
          if (b >= (size_t)NUM_BUCKETS) throw std::runtime_error("bucket index out of bounds");
          buckets[b].count += 1;
          buckets[b].b = buckets[b].b.united(p.bounds);
        }
        float costs[NUM_BUCKETS - 1];
        for (int i = 0; i < NUM_BUCKETS - 1; ++i) {
          size_t cnt0 = 0, cnt1 = 0;
          BBox b0, b1;
          for (int j = 0; j <= i; ++j) {
            cnt0 += buckets[j].count;
            b0 = b0.united(buckets[j].b);
          }
          for (int j = i + 1; j < NUM_BUCKETS; ++j) {
            cnt1 += buckets[j].count;
            b1 = b1.united(buckets[j].b);
          }
          float b0sa = (float)cnt0 * b0.surface_area();
          float b1sa = (float)cnt1 * b1.surface_area();
          float sc = 0.125f;
          float tsa = bbox.surface_area();
          costs[i] = sc * (b0sa + b1sa) / tsa;  // D20: no +0.125 traversal term
        }
        size_t min_split = 0;
        float min_cost = F32_MAX;
        for (int i = 0; i < NUM_BUCKETS - 1; ++i)
          if (costs[i] < min_cost) {
            min_split = (size_t)i;
            min_cost = costs[i];
          }
        if (max_prims < num_prims || f2usize(min_cost) < num_prims) {
          for (const Info& p : v) (bucket_for(p) <= min_split ? p1 : p2).push_back(p);
        } else {
          return make_leaf(bbox, v, ordered);
        }
        break;
      }
    }
    if (p1.empty() || p2.empty()) throw std::runtime_error("assert!(p.len() > 0) (bvh.rs:237)");
    v.clear();
    v.shrink_to_fit();
    auto left = recursive_build(std::move(p1), ordered);
    auto right = recursive_build(std::move(p2), ordered);
    auto nd = std::make_unique<BuildNode>();
    nd->leaf = false;
    nd->bounds = left->bounds.united(right->bounds);
    nd->axis = dim;
    nd->num_nodes = left->num_nodes + right->num_nodes + 1;
    nd->c1 = std::move(left);
    nd->c2 = std::move(right);
    return nd;
  }

  // bvh.rs:282-316 depth-first flatten; first child is always index+1.
  void flatten(const BuildNode& n) {
    PackedNode pn;
    pn.bounds = n.bounds;
    if (n.leaf) {
      pn.leaf = true;
      pn.offset = n.first;
      pn.count = n.n;
      nodes.push_back(pn);
      return;
    }
    pn.leaf = false;
    pn.count = (uint32_t)n.axis;
    size_t me = nodes.size();
    nodes.push_back(pn);
    flatten(*n.c1);
    nodes[me].offset = (uint32_t)nodes.size();
    flatten(*n.c2);
  }

  // bvh.rs:375-422  closest hit.  `ray.maxt` is left at the hit distance as in the reference.
  bool intersect(const Ray& ray, Hit* out, TraceCounters* tc = nullptr) const {
    if (nodes.empty()) return false;
    V3 inv_dir(1.f / ray.d.x, 1.f / ray.d.y, 1.f / ray.d.z);
    bool dir_is_neg[3] = {inv_dir.x < 0.0f, inv_dir.y < 0.0f, inv_dir.z < 0.0f};
    std::vector<uint32_t> todo;
    todo.reserve(64);
    todo.push_back(0);
    bool found = false;
    while (!todo.empty()) {
      uint32_t node_num = todo.back();
      todo.pop_back();
      if (tc) tc->nodes_visited++;
      const PackedNode& nd = nodes[node_num];
      if (!nd.bounds.intersect(ray)) continue;
      if (nd.leaf) {
        for (uint32_t i = 0; i < nd.count; ++i) {
          Hit h;
          if (prim_intersect(prims[nd.offset + i], ray, &h, tc)) {
            h.prim = nd.offset + i;
            *out = h;  // last Some wins (bvh.rs:400-405)
            found = true;
          }
        }
      } else {
        if (dir_is_neg[nd.count]) {
          todo.push_back(node_num + 1);
          todo.push_back(nd.offset);
        } else {
          todo.push_back(nd.offset);
          todo.push_back(node_num + 1);
        }
      }
    }
    return found;
  }

  // intersection.rs:63-65 default intersect_p == intersect(r).is_some().  The boolean equals that
  // of an early-exit any-hit traversal; `early_exit` selects which one is *counted* (the
  // reference does the full closest-hit walk; the roofline's shadow-ray bytes use early exit).
  bool intersect_p(const Ray& ray, bool early_exit, TraceCounters* tc = nullptr) const {
    if (!early_exit) {
      Hit h;
      return intersect(ray, &h, tc);
    }
    if (nodes.empty()) return false;
    V3 inv_dir(1.f / ray.d.x, 1.f / ray.d.y, 1.f / ray.d.z);
    bool dir_is_neg[3] = {inv_dir.x < 0.0f, inv_dir.y < 0.0f, inv_dir.z < 0.0f};
    std::vector<uint32_t> todo;
    todo.reserve(64);
    todo.push_back(0);
    while (!todo.empty()) {
      uint32_t node_num = todo.back();
      todo.pop_back();
      if (tc) tc->nodes_visited++;
      const PackedNode& nd = nodes[node_num];
      if (!nd.bounds.intersect(ray)) continue;
      if (nd.leaf) {
        for (uint32_t i = 0; i < nd.count; ++i) {
          Hit h;
          Ray probe = ray;  // do not shrink maxt: any hit in [mint,maxt] answers the query
          if (prim_intersect(prims[nd.offset + i], probe, &h, tc)) return true;
        }
      } else {
        if (dir_is_neg[nd.count]) {
          todo.push_back(node_num + 1);
          todo.push_back(nd.offset);
        } else {
          todo.push_back(nd.offset);
          todo.push_back(node_num + 1);
        }
      }
    }
    return false;
  }
};

// ---------------------------------------------------------------------------------------------
// Geometry container + refinement order (primitive/mod.rs:48-62,194-216; shape/mesh.rs:324-335;
// SURVEY Appendix A): the list handed to the BVH builder holds, per user-level primitive in input
// order, a sphere as itself or a mesh's triangles in ORIGINAL triple order j = 0..n-1, each with
// vertex order (vi[3j+2], vi[3j+1], vi[3j]).
struct Geometry {
  std::vector<std::unique_ptr<Mesh>> meshes;
  std::vector<std::unique_ptr<Sphere>> spheres;
  std::vector<Prim> refined;
  uint32_t n_objects = 0;

  void add_mesh(std::unique_ptr<Mesh> m) {
    const Mesh* mp = m.get();
    meshes.push_back(std::move(m));
    size_t nt = mp->vi.size() / 3;
    // Mesh::refine pops from the end -> tris in reverse triple order; fully_refine pops from the
    // end again -> original order.  Emulate both reversals explicitly.
    std::vector<Prim> tris;
    tris.reserve(nt);
    std::vector<uint32_t> indices = mp->vi;
    uint32_t j = (uint32_t)nt;
    while (indices.size() >= 3) {
      uint32_t v1 = indices.back();
      indices.pop_back();
      uint32_t v2 = indices.back();
      indices.pop_back();
      uint32_t v3 = indices.back();
      indices.pop_back();
      Prim p;
      p.kind = Prim::TRI;
      p.mesh = mp;
      p.v[0] = v1;
      p.v[1] = v2;
      p.v[2] = v3;
      p.source_index = --j;
      p.source_object = n_objects;
      tris.push_back(p);
    }
    while (!tris.empty()) {  // FullyRefinable::fully_refine
      refined.push_back(tris.back());
      tris.pop_back();
    }
    n_objects++;
  }
  void add_sphere(std::unique_ptr<Sphere> s) {
    Prim p;
    p.kind = Prim::SPH;
    p.sphere = s.get();
    p.source_index = 0;  // index within its user-level object
    p.source_object = n_objects++;
    spheres.push_back(std::move(s));
    refined.push_back(p);
  }
};

}  // namespace orc
