// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// extern "C" surface of the CPU oracle, loaded with ctypes by tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs (nothing else may load it).
#include <chrono>
#include <limits>
#include <cstdio>
#include <string>

#include "render.hpp"

using namespace orc;

namespace {
thread_local std::string g_err;
Transform xf_from(const float* m, const float* minv) {
  Transform t;
  std::memcpy(t.m.m, m, 64);
  std::memcpy(t.m_inv.m, minv, 64);
  return t;
}
template <class F>
int guarded(F f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
}  // namespace

struct OrcScene {
  Scene sc;
  bool built = false;
};

// Plain-data render description (mirrors the arguments of Camera::perspective, Film::image,
// Sampler::stratified / low_discrepancy and SamplerRenderer::new in the reference).
struct OrcRenderConfig {
  float cam_to_world[16];
  float cam_to_world_inv[16];
  float screen_window[4];
  float sopen, sclose, lensr, focald, fov;
  int32_t x_res, y_res;
  float crop[4];
  int32_t filter_type;
  float filter_xw, filter_yw, filter_p0, filter_p1;
  int32_t sampler_kind, xs, ys, jitter;
  int32_t num_tasks;  // 0 -> num_tasks_for(num_cpus, x_res*y_res)
  int32_t num_cpus;
  int32_t mode;  // 0 default, 1 strict
  int32_t n_threads;
  int32_t count_traversal;
  int32_t primary_only;
};

struct OrcRenderStats {
  uint64_t camera_rays, camera_hits, shadow_rays;
  uint64_t nodes_visited, tris_tested, spheres_tested;
  uint64_t sh_nodes_visited, sh_tris_tested, sh_spheres_tested;
  uint64_t nan_samples;
  double seconds;
  int32_t num_tasks;
  int32_t sample_ext[4];
  int32_t pixel_ext[4];
};

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

OrcScene* orc_scene_new() { return new OrcScene(); }
void orc_scene_free(OrcScene* s) { delete s; }

// Shape::triangle_mesh + Primitive::geometric[_area_light]
int orc_add_mesh(OrcScene* s, const float* o2w, const float* o2w_inv, int ro, const uint32_t* idx,
                 uint64_t n_idx, const float* P, uint64_t n_p, const float* N, const float* S,
                 const float* UV, uint32_t material, int32_t area_light) {
  return guarded([&] {
    Transform t = xf_from(o2w, o2w_inv);
    auto m = std::make_unique<Mesh>(t, t.inverse(), ro != 0, idx, n_idx, P, n_p, N, S, UV);
    m->material = material;
    m->area_light = area_light;
    s->sc.geom.add_mesh(std::move(m));
  });
}
// Shape::sphere(o2w, w2o, ro, rad, z0, z1, pm)
int orc_add_sphere(OrcScene* s, const float* o2w, const float* o2w_inv, int ro, float rad,
                   float z0, float z1, float pm, uint32_t material) {
  return guarded([&] {
    Transform t = xf_from(o2w, o2w_inv);
    auto sp = std::make_unique<Sphere>(t, t.inverse(), ro != 0, rad, z0, z1, pm);
    sp->material = material;
    s->sc.geom.add_sphere(std::move(sp));
  });
}
// Primitive::geometric(Shape::cylinder(o2w, w2o, ro, rad, z0, z1, phi_max), material) and
// Shape::disk(o2w, w2o, ro, height, radius, inner_radius, phi_max)   (shape/mod.rs)
int orc_add_cylinder(OrcScene* s, const float* o2w, const float* o2w_inv, int ro, float rad, float z0, float z1,
                     float pm, uint32_t material) {
  return guarded([&] {
    Transform t = xf_from(o2w, o2w_inv);
    auto sp = std::make_unique<Sphere>(Sphere::CylinderTag{}, t, t.inverse(), ro != 0, rad, z0, z1, pm);
    sp->material = material;
    s->sc.geom.add_sphere(std::move(sp));
  });
}
int orc_add_disk(OrcScene* s, const float* o2w, const float* o2w_inv, int ro, float ht, float r, float ri,
                 float pm, uint32_t material) {
  return guarded([&] {
    Transform t = xf_from(o2w, o2w_inv);
    auto sp = std::make_unique<Sphere>(Sphere::DiskTag{}, t, t.inverse(), ro != 0, ht, r, ri, pm);
    sp->material = material;
    s->sc.geom.add_sphere(std::move(sp));
  });
}
// map12: UV (su,sv,du,dv) | Planar (vs, vt, ds, dt) | Spherical / Cylindrical / Identity3D: rows
// 0..2 of world_to_texture.m (row 3 = 0 0 0 1, the affine case the device supports)
static Mapping2D mapping_from(int map_kind, const float* m) {
  Mapping2D mp;
  mp.kind = map_kind;
  if (map_kind == 0) {
    mp.su = m[0];
    mp.sv = m[1];
    mp.du = m[2];
    mp.dv = m[3];
  } else if (map_kind == 1) {
    mp.vs = V3(m[0], m[1], m[2]);
    mp.vt = V3(m[3], m[4], m[5]);
    mp.du = m[6];
    mp.dv = m[7];
  } else {
    M44 a = M44::rows(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9], m[10], m[11], 0, 0, 0, 1);
    mp.w2t = Transform(a, a);  // only m is read by the mappings
  }
  return mp;
}
int orc_add_texture(OrcScene* s, int kind, const float* value12, int map_kind, const float* map12,
                    int tex1, int tex2, int tex3, int aa) {
  Texture t;
  t.kind = kind;
  t.value = RGB(value12[0], value12[1], value12[2]);
  for (int k = 0; k < 4; ++k) t.bil[k] = RGB(value12[3 * k], value12[3 * k + 1], value12[3 * k + 2]);
  t.mapping = mapping_from(map_kind, map12);
  t.tex1 = tex1;
  t.tex2 = tex2;
  t.tex3 = tex3;
  t.aa = aa;
  s->sc.textures.t.push_back(t);
  return (int)s->sc.textures.t.size() - 1;
}
// TextureCache::new_texture (texture/imagemap.rs:128-138 / 183-193).  rgb = read_image's texels
// (byte / 255, row-major) or NULL for an unreadable file (1x1 scale^gamma map, :116-120).
int orc_add_image_texture(OrcScene* s, int map_kind, const float* map12, const float* rgb, uint32_t w,
                          uint32_t h, int spectrum, int do_trilinear, float max_aniso, int wrap,
                          float scale, float gamma) {
  const float zero[12] = {0.f};
  int id = orc_add_texture(s, 3, zero, map_kind, map12, 0, 0, 0, 0);
  s->sc.textures.mips.push_back(
      make_image_mipmap(rgb, w, h, spectrum != 0, do_trilinear != 0, max_aniso, wrap, scale, gamma));
  s->sc.textures.t[(size_t)id].tex1 = (int)s->sc.textures.mips.size() - 1;
  return id;
}
int orc_add_material(OrcScene* s, int kind, int kd, int sigma, int ks, int roughness, int bump) {
  Material m;
  m.bump = bump;
  m.kind = kind;
  m.kd = kd;
  m.sigma = sigma;
  m.ks = ks;
  m.roughness = roughness;
  s->sc.materials.push_back(m);
  return (int)s->sc.materials.size() - 1;
}
// PointLight::new(l2w, intensity) / SpotLight::new(l2w, intensity, width, fall)
int orc_add_point_light(OrcScene* s, const float* l2w, const float* l2w_inv, const float* I) {
  Transform t = xf_from(l2w, l2w_inv);
  Light l;
  l.kind = 0;
  l.pos = t.pt(V3());
  l.intensity = RGB(I[0], I[1], I[2]);
  l.world_to_light = t.inverse();
  s->sc.lights.push_back(l);
  return (int)s->sc.lights.size() - 1;
}
int orc_add_spot_light(OrcScene* s, const float* l2w, const float* l2w_inv, const float* I,
                       float width, float fall) {
  Transform t = xf_from(l2w, l2w_inv);
  Light l;
  l.kind = 1;
  l.pos = t.pt(V3());
  l.intensity = RGB(I[0], I[1], I[2]);
  l.world_to_light = t.inverse();
  l.cos_total_width = std::cos(as_radians(width));
  l.cos_falloff_start = std::cos(as_radians(fall));
  s->sc.lights.push_back(l);
  return (int)s->sc.lights.size() - 1;
}
// Oracle-defined diffuse area light; meshes added with area_light == the returned id emit.
int orc_add_area_light(OrcScene* s, const float* L, int num_samples) {
  Light l;
  l.kind = 2;
  l.intensity = RGB(L[0], L[1], L[2]);
  l.num_samples = num_samples;
  s->sc.lights.push_back(l);
  return (int)s->sc.lights.size() - 1;
}

// BVHAccelerator::new(prims, max_prims, sm) ; sm: 0 middle, 1 equal, 2 sah
int orc_build(OrcScene* s, uint32_t max_prims, int split_method) {
  return guarded([&] {
    Scene& sc = s->sc;
    sc.bvh.build(sc.geom.refined, max_prims, (SplitMethod)split_method);
    // collect emissive triangles per area light, in BVH-input (refined) order
    for (Light& l : sc.lights) {
      l.tris.clear();
      l.cdf.clear();
    }
    for (const Prim& p : sc.geom.refined) {
      if (p.kind != Prim::TRI || p.mesh->area_light < 0) continue;
      Light& l = sc.lights.at((size_t)p.mesh->area_light);
      if (l.kind != 2) throw std::runtime_error("mesh refers to a non-area light");
      Light::Tri t;
      t.p1 = p.mesh->p[p.v[0]];
      t.p2 = p.mesh->p[p.v[1]];
      t.p3 = p.mesh->p[p.v[2]];
      Ray dummy(V3(), V3(0, 0, 1), 0.f);
      DiffGeom dg = tri_dg(p, dummy, 0.f, 0.f, 0.f);
      t.nn = dg.nn;
      t.area = 0.5f * length(cross(t.p2 - t.p1, t.p3 - t.p1));  // mesh.rs:100-103
      l.tris.push_back(t);
    }
    for (Light& l : sc.lights) {
      if (l.kind != 2) continue;
      if (l.tris.empty()) throw std::runtime_error("area light without emissive triangles");
      float total = 0.f;
      for (const auto& t : l.tris) total += t.area;
      l.total_area = total;
      l.cdf.resize(l.tris.size() + 1);
      l.cdf[0] = 0.f;
      float acc = 0.f;
      for (size_t i = 0; i < l.tris.size(); ++i) {
        acc += l.tris[i].area;
        l.cdf[i + 1] = acc / total;
      }
      l.cdf.back() = 1.0f;
    }
    s->built = true;
  });
}

uint64_t orc_num_nodes(const OrcScene* s) { return s->sc.bvh.nodes.size(); }
uint64_t orc_num_prims(const OrcScene* s) { return s->sc.bvh.prims.size(); }
// bounds: 6 floats/node ; meta: (offset, count_or_axis, is_leaf) u32 triples
void orc_get_nodes(const OrcScene* s, float* bounds, uint32_t* meta) {
  const auto& n = s->sc.bvh.nodes;
  for (size_t i = 0; i < n.size(); ++i) {
    bounds[6 * i + 0] = n[i].bounds.p_min.x;
    bounds[6 * i + 1] = n[i].bounds.p_min.y;
    bounds[6 * i + 2] = n[i].bounds.p_min.z;
    bounds[6 * i + 3] = n[i].bounds.p_max.x;
    bounds[6 * i + 4] = n[i].bounds.p_max.y;
    bounds[6 * i + 5] = n[i].bounds.p_max.z;
    meta[3 * i + 0] = n[i].offset;
    meta[3 * i + 1] = n[i].count;
    meta[3 * i + 2] = n[i].leaf ? 1u : 0u;
  }
}
// per ordered primitive: (kind, source_object, source_index)
void orc_get_prim_order(const OrcScene* s, uint32_t* out) {
  const auto& p = s->sc.bvh.prims;
  for (size_t i = 0; i < p.size(); ++i) {
    out[3 * i + 0] = (uint32_t)p[i].kind;
    out[3 * i + 1] = p[i].source_object;
    out[3 * i + 2] = p[i].source_index;
  }
}

// rays: 8 floats each (ox,oy,oz,mint, dx,dy,dz,maxt).  hits: (prim u32, t, b1, b2) as 4x u32/f32.
// counters (optional): 3 u32 per ray (nodes, tris, spheres).  maxt_out (optional): final ray.maxt.
int orc_trace_closest(const OrcScene* s, const float* rays, uint64_t n, uint32_t* hit_prim,
                      float* hit_tbb, uint32_t* counters, int n_threads) {
  return guarded([&] {
    auto work = [&](uint64_t lo, uint64_t hi) {
      for (uint64_t i = lo; i < hi; ++i) {
        const float* r = rays + 8 * i;
        Ray ray(V3(r[0], r[1], r[2]), V3(r[4], r[5], r[6]), r[3]);
        ray.maxt = r[7];
        Hit h;
        TraceCounters tc;
        bool f = s->sc.bvh.intersect(ray, &h, counters ? &tc : nullptr);
        hit_prim[i] = f ? h.prim : 0xFFFFFFFFu;
        hit_tbb[3 * i + 0] = f ? h.t : 0.f;
        hit_tbb[3 * i + 1] = f ? h.b1 : 0.f;
        hit_tbb[3 * i + 2] = f ? h.b2 : 0.f;
        if (counters) {
          counters[3 * i + 0] = (uint32_t)tc.nodes_visited;
          counters[3 * i + 1] = (uint32_t)tc.tris_tested;
          counters[3 * i + 2] = (uint32_t)tc.spheres_tested;
        }
      }
    };
    int nt = std::max(1, n_threads);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back(work, n * t / nt, n * (t + 1) / nt);
    for (auto& t : th) t.join();
  });
}
// occluded[i] = 1 iff any hit in [mint, maxt]  (== Scene::intersect_p)
int orc_trace_any(const OrcScene* s, const float* rays, uint64_t n, uint8_t* occluded,
                  uint32_t* counters, int early_exit, int n_threads) {
  return guarded([&] {
    auto work = [&](uint64_t lo, uint64_t hi) {
      for (uint64_t i = lo; i < hi; ++i) {
        const float* r = rays + 8 * i;
        Ray ray(V3(r[0], r[1], r[2]), V3(r[4], r[5], r[6]), r[3]);
        ray.maxt = r[7];
        TraceCounters tc;
        occluded[i] = s->sc.bvh.intersect_p(ray, early_exit != 0, counters ? &tc : nullptr) ? 1 : 0;
        if (counters) {
          counters[3 * i + 0] = (uint32_t)tc.nodes_visited;
          counters[3 * i + 1] = (uint32_t)tc.tris_tested;
          counters[3 * i + 2] = (uint32_t)tc.spheres_tested;
        }
      }
    };
    int nt = std::max(1, n_threads);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back(work, n * t / nt, n * (t + 1) / nt);
    for (auto& t : th) t.join();
  });
}

static RenderConfig make_config(const OrcScene* s, const OrcRenderConfig* c) {
  RenderConfig cfg;
  Transform c2w = xf_from(c->cam_to_world, c->cam_to_world_inv);
  cfg.camera = PerspectiveCamera(c2w, c->screen_window, c->sopen, c->sclose, c->lensr, c->focald,
                                 c->fov, c->x_res, c->y_res);
  Filter f(c->filter_type, c->filter_xw, c->filter_yw, c->filter_p0, c->filter_p1);
  cfg.film = Film(c->x_res, c->y_res, f, c->crop);
  cfg.sampler.kind = c->sampler_kind;
  cfg.film.sample_extent(cfg.sampler.ext);  // Sampler built from film.get_sample_extent()
  cfg.sampler.xs = c->xs;
  cfg.sampler.ys = c->ys;
  cfg.sampler.jitter = c->jitter != 0;
  cfg.sampler.sopen = c->sopen;
  cfg.sampler.sclose = c->sclose;
  cfg.sampler.light_samples = s ? s->sc.light_sample_pairs() : 0;
  cfg.num_tasks = c->num_tasks > 0
                      ? (uint32_t)c->num_tasks
                      : num_tasks_for((uint32_t)std::max(1, c->num_cpus),
                                      (uint32_t)(c->x_res * c->y_res));
  cfg.mode = c->mode;
  cfg.n_threads = c->n_threads;
  cfg.count_traversal = c->count_traversal != 0;
  cfg.primary_only = c->primary_only != 0;
  return cfg;
}

// Fills stats->sample_ext / pixel_ext / num_tasks without rendering.
int orc_render_layout(const OrcRenderConfig* c, OrcRenderStats* stats) {
  return guarded([&] {
    RenderConfig cfg = make_config(nullptr, c);
    std::memset(stats, 0, sizeof *stats);
    stats->num_tasks = (int32_t)cfg.num_tasks;
    for (int i = 0; i < 4; ++i) stats->sample_ext[i] = cfg.sampler.ext[i];
    cfg.film.pixel_extent(stats->pixel_ext);
  });
}

// film_xyzw: 4 floats per film pixel (pixel extent, row-major): xyz sums + weight_sum.
// rgb (optional): 3 floats per pixel, D6 conversion.  hit_ids/hit_ts (optional): per camera sample
// over the sampler extent.
int orc_render(OrcScene* s, const OrcRenderConfig* c, float* film_xyzw, float* rgb,
               uint32_t* hit_ids, float* hit_ts, OrcRenderStats* stats) {
  return guarded([&] {
    if (!s->built) throw std::runtime_error("scene not built");
    RenderConfig cfg = make_config(s, c);
    RenderStats st;
    auto t0 = std::chrono::steady_clock::now();
    render(s->sc, cfg, &st, hit_ids, hit_ts);
    auto t1 = std::chrono::steady_clock::now();
    if (film_xyzw)
      for (size_t i = 0; i < cfg.film.pixels.size(); ++i) {
        film_xyzw[4 * i + 0] = cfg.film.pixels[i].xyz[0];
        film_xyzw[4 * i + 1] = cfg.film.pixels[i].xyz[1];
        film_xyzw[4 * i + 2] = cfg.film.pixels[i].xyz[2];
        film_xyzw[4 * i + 3] = cfg.film.pixels[i].weight_sum;
      }
    if (rgb) cfg.film.to_rgb(rgb);
    if (stats) {
      std::memset(stats, 0, sizeof *stats);
      stats->camera_rays = st.camera_rays;
      stats->camera_hits = st.camera_hits;
      stats->shadow_rays = st.shadow_rays;
      stats->nodes_visited = st.nodes_visited;
      stats->tris_tested = st.tris_tested;
      stats->spheres_tested = st.spheres_tested;
      stats->sh_nodes_visited = st.sh_nodes_visited;
      stats->sh_tris_tested = st.sh_tris_tested;
      stats->sh_spheres_tested = st.sh_spheres_tested;
      stats->nan_samples = st.nan_samples;
      stats->seconds = std::chrono::duration<double>(t1 - t0).count();
      stats->num_tasks = (int32_t)cfg.num_tasks;
      for (int i = 0; i < 4; ++i) stats->sample_ext[i] = cfg.sampler.ext[i];
      cfg.film.pixel_extent(stats->pixel_ext);
    }
    if (st.nan_samples) throw std::runtime_error("Invalid radiance value!");  // D4
  });
}
void orc_set_strict_flags(OrcScene* s, int on) { s->sc.strict_flags = on != 0; }

// Camera samples + rays for the pixels of one row range (unit-level checks of A1-A3).
// out_cs: 5 floats per sample; out_rays: 8 floats per sample (o, mint, d, maxt);
// out_diff (optional): 12 floats (rx_origin, ry_origin, rx_dir, ry_dir) after scale_differentials.
int orc_camera_samples(const OrcRenderConfig* c, int light_pairs, int x0, int x1, int y0, int y1,
                       float* out_cs, float* out_rays, float* out_diff, float* out_light_u) {
  return guarded([&] {
    RenderConfig cfg = make_config(nullptr, c);
    cfg.sampler.light_samples = light_pairs;
    const SamplerDesc& sd = cfg.sampler;
    auto tw = task_windows(sd, cfg.num_tasks);
    size_t spp = sd.spp(), W = sd.words_per_pixel();
    std::vector<CameraSample> cs;
    std::vector<float> lu;
    size_t k = 0;
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x) {
        const TaskWindow* w = nullptr;
        for (auto& t : tw)
          if (!t.empty && x >= t.ext[0] && x < t.ext[1] && y >= t.ext[2] && y < t.ext[3]) w = &t;
        if (!w) throw std::runtime_error("pixel outside every task window");
        RNG rng = RNG::from_key(w->key);
        size_t pk = (size_t)(y - w->ext[2]) * (size_t)(w->ext[1] - w->ext[0]) + (size_t)(x - w->ext[0]);
        rng.seek(pk * W);
        pixel_samples(sd, x, y, rng, cs, lu);
        if (rng.tell() != (pk + 1) * W) throw std::runtime_error("words_per_pixel mismatch");
        for (size_t i = 0; i < spp; ++i, ++k) {
          float* o = out_cs + 5 * k;
          o[0] = cs[i].image_x;
          o[1] = cs[i].image_y;
          o[2] = cs[i].lens_u;
          o[3] = cs[i].lens_v;
          o[4] = cs[i].time;
          RayDifferential rd = cfg.camera.generate_ray_differential(cs[i]);
          rd.scale_differentials(1.0f / std::sqrt((float)spp));
          if (out_rays) {
            float* r = out_rays + 8 * k;
            r[0] = rd.ray.o.x; r[1] = rd.ray.o.y; r[2] = rd.ray.o.z; r[3] = rd.ray.mint;
            r[4] = rd.ray.d.x; r[5] = rd.ray.d.y; r[6] = rd.ray.d.z; r[7] = rd.ray.maxt;
          }
          if (out_diff) {
            float* d = out_diff + 12 * k;
            const V3* v[4] = {&rd.rx_origin, &rd.ry_origin, &rd.rx_dir, &rd.ry_dir};
            for (int q = 0; q < 4; ++q) {
              d[3 * q] = v[q]->x; d[3 * q + 1] = v[q]->y; d[3 * q + 2] = v[q]->z;
            }
          }
          if (out_light_u)
            for (int q = 0; q < 2 * light_pairs; ++q)
              out_light_u[(size_t)(2 * light_pairs) * k + q] = lu[(size_t)(2 * light_pairs) * i + q];
        }
      }
  });
}

// HaltonSampler: slots per pixel of the padded per-sample layout (the largest per-pixel count)
// and, optionally, the per-pixel counts over the full sampler extent.
int orc_halton_cap(const OrcRenderConfig* c, int light_pairs, uint32_t* out_cap, uint32_t* out_counts) {
  return guarded([&] {
    RenderConfig cfg = make_config(nullptr, c);
    cfg.sampler.light_samples = light_pairs;
    HaltonLayout L = halton_layout(cfg.sampler, cfg.num_tasks, std::max(1, cfg.n_threads));
    *out_cap = L.cap;
    if (out_counts) std::memcpy(out_counts, L.counts.data(), L.counts.size() * sizeof(uint32_t));
  });
}
// HaltonSampler camera samples in the padded layout [pixel raster over the full extent][cap]:
// out_cs 5 floats (NaN image coordinates mark unused slots), out_light_u 2 * light_pairs floats.
int orc_halton_samples(const OrcRenderConfig* c, int light_pairs, uint32_t cap, float* out_cs, float* out_light_u) {
  return guarded([&] {
    RenderConfig cfg = make_config(nullptr, c);
    cfg.sampler.light_samples = light_pairs;
    HaltonLayout L = halton_layout(cfg.sampler, cfg.num_tasks, std::max(1, cfg.n_threads));
    if (cap != L.cap) throw std::runtime_error("halton: cap mismatch");
    const size_t total = L.counts.size() * (size_t)cap, lf = 2 * (size_t)light_pairs;
    const float nan = std::numeric_limits<float>::quiet_NaN();
    for (size_t k = 0; k < total; ++k) {
      out_cs[5 * k] = out_cs[5 * k + 1] = nan;
      out_cs[5 * k + 2] = out_cs[5 * k + 3] = out_cs[5 * k + 4] = 0.f;
    }
    for (size_t k = 0; k < L.entries.size(); ++k) {
      const HaltonEntry& e = L.entries[k];
      const size_t o = (size_t)e.pixel * cap + e.slot;
      out_cs[5 * o] = e.cs.image_x;
      out_cs[5 * o + 1] = e.cs.image_y;
      out_cs[5 * o + 2] = e.cs.lens_u;
      out_cs[5 * o + 3] = e.cs.lens_v;
      out_cs[5 * o + 4] = e.cs.time;
      if (out_light_u)
        for (size_t q = 0; q < lf; ++q) out_light_u[lf * o + q] = L.light_u[lf * k + q];
    }
  });
}
double orc_radical_inverse(uint64_t n, uint64_t b) { return radical_inverse(n, b); }

// ---- shading-stage hooks (used by the device-source checks, tests/test_device_source.py) --------
// BSDF::f (bsdf/mod.rs:132-149) for one material evaluated to constants.
// kind: 0 matte/Lambertian, 1 matte/OrenNayar (sigma in degrees), 2 plastic (Lambertian + Blinn, roughness).
// frame9 = shading nn, geometric ng, dpdu (BSDF::new_with_eta builds sn, tn from nn and dpdu).
void orc_bsdf_f(int kind, const float* kd3, const float* ks3, float sigma_or_rough, const float* frame9,
                const float* wo3, const float* wi3, int strict_flags, float* out3) {
  DiffGeom dgs;
  dgs.nn = V3(frame9[0], frame9[1], frame9[2]);
  dgs.dpdu = V3(frame9[6], frame9[7], frame9[8]);
  BSDF bsdf(dgs, V3(frame9[3], frame9[4], frame9[5]));
  BxDF d;
  d.r = RGB(kd3[0], kd3[1], kd3[2]);
  if (kind == 1) {  // orennayar.rs:16-29
    d.kind = 1;
    float sigma = as_radians(sigma_or_rough);
    float sigma2 = sigma * sigma;
    d.a = 1.0f - (sigma2 / (2.0f * (sigma + 0.33f)));
    d.b = 0.45f * sigma2 / (sigma2 + 0.09f);
  } else {
    d.kind = 0;
  }
  bsdf.bxdfs[bsdf.n_bxdfs++] = d;
  if (kind == 2) {
    BxDF sp;
    sp.kind = 2;
    sp.r = RGB(ks3[0], ks3[1], ks3[2]);
    float e = 1.0f / sigma_or_rough;
    if (e > 1000.0f || std::isnan(e)) e = 1000.0f;  // microfacet.rs:18-24
    sp.a = e;
    bsdf.bxdfs[bsdf.n_bxdfs++] = sp;
  }
  RGB f = bsdf.f(V3(wo3[0], wo3[1], wo3[2]), V3(wi3[0], wi3[1], wi3[2]), strict_flags != 0);
  out3[0] = f.c[0]; out3[1] = f.c[1]; out3[2] = f.c[2];
}
float orc_fresnel_dielectric(float cosi, float eta_i, float eta_t) { return BxDF::fresnel_dielectric(cosi, eta_i, eta_t); }
// DifferentialGeometry::compute_differentials (diff_geom.rs:81-152).  g12 = p, nn, dpdu, dpdv;
// rd12 = rx_origin, ry_origin, rx_dir, ry_dir; out10 = dpdx, dpdy, dudx, dvdx, dudy, dvdy.
void orc_compute_differentials(const float* g12, const float* rd12, float* out10) {
  DiffGeom dg;
  dg.p = V3(g12[0], g12[1], g12[2]);
  dg.nn = V3(g12[3], g12[4], g12[5]);
  dg.dpdu = V3(g12[6], g12[7], g12[8]);
  dg.dpdv = V3(g12[9], g12[10], g12[11]);
  RayDifferential rd;
  rd.has_differentials = true;
  rd.rx_origin = V3(rd12[0], rd12[1], rd12[2]);
  rd.ry_origin = V3(rd12[3], rd12[4], rd12[5]);
  rd.rx_dir = V3(rd12[6], rd12[7], rd12[8]);
  rd.ry_dir = V3(rd12[9], rd12[10], rd12[11]);
  dg.compute_differentials(rd);
  out10[0] = dg.dpdx.x; out10[1] = dg.dpdx.y; out10[2] = dg.dpdx.z;
  out10[3] = dg.dpdy.x; out10[4] = dg.dpdy.y; out10[5] = dg.dpdy.z;
  out10[6] = dg.dudx; out10[7] = dg.dvdx; out10[8] = dg.dudy; out10[9] = dg.dvdy;
}
// VisibilityTester::segment (visibility_tester.rs:16-24): ray8 = o, mint, d, maxt
void orc_vis_segment(const float* p1, float eps1, const float* p2, float eps2, float* ray8) {
  Ray r = vis_segment(V3(p1[0], p1[1], p1[2]), eps1, V3(p2[0], p2[1], p2[2]), eps2, 0.f);
  ray8[0] = r.o.x; ray8[1] = r.o.y; ray8[2] = r.o.z; ray8[3] = r.mint;
  ray8[4] = r.d.x; ray8[5] = r.d.y; ray8[6] = r.d.z; ray8[7] = r.maxt;
}

// Triangle::intersect's dg part + get_shading_geometry (mesh.rs:105-193, 220-262) for ONE triangle.
// P9 = object-space p1, p2, p3 in the order the Triangle holds them (refine-reversed); N9 / S9 / UV6
// per-vertex normals / tangents / uvs or NULL.  out_dg14 = p, nn, u, v, dpdu, dpdv (geometric);
// out_dgs17 = nn, dpdu, dpdv, dndu, dndv, u, v (shading).
int orc_tri_surface(const float* o2w, const float* o2w_inv, int ro, const float* P9, const float* N9,
                    const float* S9, const float* UV6, const float* ray8, float t, float b1, float b2,
                    float* out_dg14, float* out_dgs17) {
  return guarded([&] {
    Transform xf = xf_from(o2w, o2w_inv);
    const uint32_t idx[3] = {0, 1, 2};
    Mesh m(xf, xf.inverse(), ro != 0, idx, 3, P9, 3, N9, S9, UV6);
    Prim pr;
    pr.kind = Prim::TRI;
    pr.mesh = &m;
    pr.v[0] = 0; pr.v[1] = 1; pr.v[2] = 2;
    Ray ray(V3(ray8[0], ray8[1], ray8[2]), V3(ray8[4], ray8[5], ray8[6]), ray8[3]);
    ray.maxt = ray8[7];
    DiffGeom dg = tri_dg(pr, ray, t, b1, b2);
    DiffGeom dgs = tri_shading_geometry(pr, dg);
    const float a[14] = {dg.p.x, dg.p.y, dg.p.z, dg.nn.x, dg.nn.y, dg.nn.z, dg.u, dg.v,
                         dg.dpdu.x, dg.dpdu.y, dg.dpdu.z, dg.dpdv.x, dg.dpdv.y, dg.dpdv.z};
    std::memcpy(out_dg14, a, sizeof a);
    const float b[17] = {dgs.nn.x, dgs.nn.y, dgs.nn.z, dgs.dpdu.x, dgs.dpdu.y, dgs.dpdu.z, dgs.dpdv.x, dgs.dpdv.y, dgs.dpdv.z,
                         dgs.dndu.x, dgs.dndu.y, dgs.dndu.z, dgs.dndv.x, dgs.dndv.y, dgs.dndv.z, dgs.u, dgs.v};
    std::memcpy(out_dgs17, b, sizeof b);
  });
}
// Film::add_sample over n samples in the given order: cs2 = image (x, y) pairs, rgb3 = radiance.
// out_xyzw = xyz sums + weight_sum per film pixel (pixel extent, row-major).
int orc_film_accumulate(const OrcRenderConfig* c, const float* cs2, const float* rgb3, uint64_t n, float* out_xyzw) {
  return guarded([&] {
    RenderConfig cfg = make_config(nullptr, c);
    for (uint64_t i = 0; i < n; ++i) {
      CameraSample cs;
      cs.image_x = cs2[2 * i];
      cs.image_y = cs2[2 * i + 1];
      cfg.film.add_sample(cs, rgb3 + 3 * i);
    }
    for (size_t i = 0; i < cfg.film.pixels.size(); ++i) {
      out_xyzw[4 * i + 0] = cfg.film.pixels[i].xyz[0];
      out_xyzw[4 * i + 1] = cfg.film.pixels[i].xyz[1];
      out_xyzw[4 * i + 2] = cfg.film.pixels[i].xyz[2];
      out_xyzw[4 * i + 3] = cfg.film.pixels[i].weight_sum;
    }
  });
}
// Film::add_sample (film.rs:192-249) of ONE sample with L = (1, 1, 1) into an empty film:
// out_w = weight_sum per film pixel (pixel extent, row-major).
int orc_film_add_sample(const OrcRenderConfig* c, float image_x, float image_y, float* out_w) {
  return guarded([&] {
    RenderConfig cfg = make_config(nullptr, c);
    CameraSample cs;
    cs.image_x = image_x;
    cs.image_y = image_y;
    const float rgb[3] = {1.f, 1.f, 1.f};
    cfg.film.add_sample(cs, rgb);
    for (size_t i = 0; i < cfg.film.pixels.size(); ++i) out_w[i] = cfg.film.pixels[i].weight_sum;
  });
}

// ---- small known-answer hooks (each mirrors one reference function) ----
int orc_quadratic(float a, float b, float c, float* t0, float* t1) {
  return quadratic(a, b, c, t0, t1) ? 1 : 0;
}
int orc_solve_2x2(const float* a4, const float* b2, float* x) {
  float a[2][2] = {{a4[0], a4[1]}, {a4[2], a4[3]}};
  return solve_linear_system_2x2(a, b2, &x[0], &x[1]) ? 1 : 0;
}
void orc_partition_by_i32(int32_t* v, uint64_t n) {
  partition_by(
      0, (size_t)n, [&](size_t i) { return v[i]; }, [&](size_t i, size_t j) { std::swap(v[i], v[j]); });
}
void orc_get_crop_window(uint64_t num, uint64_t count, float aspect, float* out4) {
  get_crop_window(num, count, aspect, out4);
}
void orc_compute_sub_window(const int32_t* ext4, uint64_t num, uint64_t count, int32_t* out4) {
  int e[4] = {ext4[0], ext4[1], ext4[2], ext4[3]}, o[4];
  compute_sub_window(e, num, count, o);
  for (int i = 0; i < 4; ++i) out4[i] = o[i];
}
uint32_t orc_num_tasks_for(uint32_t ncpu, uint32_t npix) { return num_tasks_for(ncpu, npix); }
// BBox::intersect: box = 6 floats, ray = 8 floats. returns 1 and (t0,t1) on hit.
int orc_bbox_intersect(const float* box, const float* r, float* t01) {
  BBox b(V3(box[0], box[1], box[2]), V3(box[3], box[4], box[5]));
  Ray ray(V3(r[0], r[1], r[2]), V3(r[4], r[5], r[6]), r[3]);
  ray.maxt = r[7];
  return b.intersect(ray, &t01[0], &t01[1]) ? 1 : 0;
}
void orc_bbox_props(const float* box, float* area, float* volume, int32_t* max_extent,
                    int32_t* empty) {
  BBox b(V3(box[0], box[1], box[2]), V3(box[3], box[4], box[5]));
  *area = b.surface_area();
  *volume = b.volume();
  *max_extent = b.max_extent();
  *empty = b.empty() ? 1 : 0;
}
int orc_tri_intersect(const float* p9, const float* r, float* tbb) {
  Ray ray(V3(r[0], r[1], r[2]), V3(r[4], r[5], r[6]), r[3]);
  ray.maxt = r[7];
  return tri_intersection_point(V3(p9[0], p9[1], p9[2]), V3(p9[3], p9[4], p9[5]),
                                V3(p9[6], p9[7], p9[8]), ray, &tbb[0], &tbb[1], &tbb[2])
             ? 1
             : 0;
}
// Sphere::intersect: returns 1 + (t_hit, ray_epsilon, phi) and the dg (p,nn,u,v,dpdu,dpdv = 14 f)
int orc_sphere_intersect(const float* o2w, const float* o2w_inv, int ro, float rad, float z0,
                         float z1, float pm, const float* r, float* out3, float* dg14) {
  Transform t = xf_from(o2w, o2w_inv);
  Sphere s(t, t.inverse(), ro != 0, rad, z0, z1, pm);
  Ray ray(V3(r[0], r[1], r[2]), V3(r[4], r[5], r[6]), r[3]);
  ray.maxt = r[7];
  Ray oray = xf_ray(s.base.w2o, ray);
  float th, phi;
  if (!s.intersection_point(oray, &th, &phi)) return 0;
  out3[0] = th;
  out3[1] = th * 5e-4f;
  out3[2] = phi;
  if (dg14) {
    DiffGeom dg = sphere_dg(s, ray, th, phi);
    float v[14] = {dg.p.x, dg.p.y, dg.p.z, dg.nn.x, dg.nn.y, dg.nn.z, dg.u, dg.v,
                   dg.dpdu.x, dg.dpdu.y, dg.dpdu.z, dg.dpdv.x, dg.dpdv.y, dg.dpdv.z};
    std::memcpy(dg14, v, sizeof v);
  }
  return 1;
}
// Known-answer hook for Cylinder (shape 1: a, b, c = rad, z0, z1) and Disk (shape 2: a, b, c =
// height, radius, inner_radius): out3 = t_hit, ray_epsilon, phi; dg14 as orc_sphere_intersect;
// props13 = radius, z_min, z_max, phi_max, height, inner_radius, area, object bound min xyz, max xyz.
int orc_quadric_intersect(int shape, const float* o2w, const float* o2w_inv, int ro, float a, float b, float c,
                          float pm, const float* r, float* out3, float* dg14, float* props13) {
  Transform t = xf_from(o2w, o2w_inv);
  Sphere s = shape == 1 ? Sphere(Sphere::CylinderTag{}, t, t.inverse(), ro != 0, a, b, c, pm)
                        : Sphere(Sphere::DiskTag{}, t, t.inverse(), ro != 0, a, b, c, pm);
  if (props13) {
    BBox ob = s.object_bound();
    float v[13] = {s.radius, s.z_min, s.z_max, s.phi_max, s.height, s.inner_radius, s.area(),
                   ob.p_min.x, ob.p_min.y, ob.p_min.z, ob.p_max.x, ob.p_max.y, ob.p_max.z};
    std::memcpy(props13, v, sizeof v);
  }
  if (!r) return 0;
  Ray ray(V3(r[0], r[1], r[2]), V3(r[4], r[5], r[6]), r[3]);
  ray.maxt = r[7];
  Ray oray = xf_ray(s.base.w2o, ray);
  float th, phi;
  if (!s.intersection_point(oray, &th, &phi)) return 0;
  out3[0] = th;
  out3[1] = th * 5e-4f;
  out3[2] = phi;
  if (dg14) {
    DiffGeom dg = sphere_dg(s, ray, th, phi);
    float v[14] = {dg.p.x, dg.p.y, dg.p.z, dg.nn.x, dg.nn.y, dg.nn.z, dg.u, dg.v,
                   dg.dpdu.x, dg.dpdu.y, dg.dpdu.z, dg.dpdv.x, dg.dpdv.y, dg.dpdv.z};
    std::memcpy(dg14, v, sizeof v);
  }
  return 1;
}
void orc_sphere_props(float rad, float z0, float z1, float pm, float* out6) {
  Sphere s(Transform(), Transform(), false, rad, z0, z1, pm);
  out6[0] = s.z_min;
  out6[1] = s.z_max;
  out6[2] = s.theta_min;
  out6[3] = s.theta_max;
  out6[4] = s.phi_max;
  out6[5] = s.phi_max * s.radius * (s.z_max - s.z_min);  // area, sphere.rs:119-121
}
int orc_invert(const float* m16, float* out16) {
  return guarded([&] {
    M44 a;
    std::memcpy(a.m, m16, 64);
    M44 r = invert(a);
    std::memcpy(out16, r.m, 64);
  });
}
void orc_look_at(const float* pos, const float* look, const float* up, float* m16, float* minv16) {
  Transform t = Transform::look_at(V3(pos[0], pos[1], pos[2]), V3(look[0], look[1], look[2]),
                                   V3(up[0], up[1], up[2]));
  std::memcpy(m16, t.m.m, 64);
  std::memcpy(minv16, t.m_inv.m, 64);
}
// Projection::new matrices for a perspective camera: raster_to_camera (m, m_inv), dx, dy.
int orc_perspective(const OrcRenderConfig* c, float* r2c16, float* r2c_inv16, float* dxdy6) {
  return guarded([&] {
    RenderConfig cfg = make_config(nullptr, c);
    std::memcpy(r2c16, cfg.camera.raster_to_camera.m.m, 64);
    std::memcpy(r2c_inv16, cfg.camera.raster_to_camera.m_inv.m, 64);
    dxdy6[0] = cfg.camera.dx_camera.x;
    dxdy6[1] = cfg.camera.dx_camera.y;
    dxdy6[2] = cfg.camera.dx_camera.z;
    dxdy6[3] = cfg.camera.dy_camera.x;
    dxdy6[4] = cfg.camera.dy_camera.y;
    dxdy6[5] = cfg.camera.dy_camera.z;
  });
}
float orc_filter_eval(int type, float xw, float yw, float p0, float p1, float x, float y) {
  return Filter(type, xw, yw, p0, p1).evaluate(x, y);
}
void orc_filter_table(int type, float xw, float yw, float p0, float p1, float* out256) {
  float crop[4] = {0, 1, 0, 1};
  Film f(4, 4, Filter(type, xw, yw, p0, p1), crop);
  std::memcpy(out256, f.table, sizeof f.table);
}
void orc_film_extents(int xres, int yres, float xw, float yw, const float* crop, int32_t* sample4,
                      int32_t* pixel4) {
  Film f(xres, yres, Filter(0, xw, yw, 0, 0), crop);
  int a[4], b[4];
  f.sample_extent(a);
  f.pixel_extent(b);
  for (int i = 0; i < 4; ++i) {
    sample4[i] = a[i];
    pixel4[i] = b[i];
  }
}
void orc_chacha_block(const uint32_t* in16, int rounds, uint32_t* out16) {
  chacha_block(in16, rounds, out16);
}
void orc_seed_from_u64(uint64_t seed, uint32_t* key8) { seed_from_u64(seed, key8); }
// first n stream words of StdRng::from_seed(key)
void orc_stream_words(const uint32_t* key8, uint64_t start, uint64_t n, uint32_t* out) {
  RNG r = RNG::from_key(key8);
  r.seek(start);
  for (uint64_t i = 0; i < n; ++i) out[i] = r.s.next_u32();
}
void orc_rng_floats(uint64_t seed, uint64_t n, float* out) {
  RNG r(seed);
  for (uint64_t i = 0; i < n; ++i) out[i] = r.random_float();
}
void orc_rng_shuffle(uint64_t seed, float* v, uint64_t len, uint64_t dims) {
  RNG r(seed);
  r.shuffle(v, len, dims);
}
void orc_stratified_2d(uint64_t seed, uint64_t nx, uint64_t ny, int jitter, float* out) {
  RNG r(seed);
  stratified_sample_2d(out, nx, ny, r, jitter != 0);
}
void orc_stratified_1d(uint64_t seed, uint64_t n, int jitter, float* out) {
  RNG r(seed);
  stratified_sample_1d(out, n, r, jitter != 0);
}
void orc_latin_hypercube(uint64_t seed, uint64_t num, uint64_t dim, float* out) {
  RNG r(seed);
  latin_hypercube(out, num, dim, r);
}
float orc_van_der_corput(uint32_t n, uint32_t scramble) { return van_der_corput(n, scramble); }
float orc_sobol2(uint32_t n, uint32_t scramble) { return sobol2(n, scramble); }

}  // extern "C"

// ---- additional known-answer hooks ----
extern "C" {
// Projection::new(film(xres,yres), proj, screen_window) -> raster_to_screen.m, screen_to_raster.m,
// raster_to_camera.m (camera/projective.rs:48-72)
int orc_projection(int xres, int yres, const float* proj, const float* proj_inv, const float* sw,
                   float* r2s16, float* s2r16, float* r2c16) {
  return guarded([&] {
    Transform p = xf_from(proj, proj_inv);
    Transform screen_to_raster = Transform::scale((float)xres, (float)yres, 1.0f) *
                                 Transform::scale(1.0f / (sw[1] - sw[0]), 1.0f / (sw[2] - sw[3]), 1.0f) *
                                 Transform::translate(V3(-sw[0], -sw[3], 0.0f));
    Transform raster_to_screen = screen_to_raster.inverse();
    Transform raster_to_cam = p.inverse() * raster_to_screen;
    std::memcpy(r2s16, raster_to_screen.m.m, 64);
    std::memcpy(s2r16, screen_to_raster.m.m, 64);
    std::memcpy(r2c16, raster_to_cam.m.m, 64);
  });
}
// Mesh::refine order and vertex triples (shape/mesh.rs:324-335): out[3*k..] = tris[k].v
void orc_mesh_refine(const uint32_t* vi, uint64_t n_vi, uint32_t* out) {
  std::vector<uint32_t> idx(vi, vi + n_vi);
  size_t k = 0;
  while (idx.size() >= 3) {
    out[3 * k + 0] = idx.back(); idx.pop_back();
    out[3 * k + 1] = idx.back(); idx.pop_back();
    out[3 * k + 2] = idx.back(); idx.pop_back();
    ++k;
  }
}
float orc_tri_area(const float* p9) {  // mesh.rs:100-103
  V3 p1(p9[0], p9[1], p9[2]), p2(p9[3], p9[4], p9[5]), p3(p9[6], p9[7], p9[8]);
  return 0.5f * length(cross(p2 - p1, p3 - p1));
}
void orc_xf_apply(const float* m, const float* minv, int kind, const float* v3, float* out3) {
  Transform t = xf_from(m, minv);
  V3 v(v3[0], v3[1], v3[2]);
  V3 r = kind == 0 ? t.pt(v) : (kind == 1 ? t.vec(v) : t.nrm(v));
  out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}
void orc_transform(int kind, const float* a3, float* m16, float* minv16) {
  Transform t;
  switch (kind) {
    case 0: t = Transform::translate(V3(a3[0], a3[1], a3[2])); break;
    case 1: t = Transform::scale(a3[0], a3[1], a3[2]); break;
    case 2: t = Transform::rotate_x(a3[0]); break;
    case 3: t = Transform::rotate_y(a3[0]); break;
    default: t = Transform::rotate_z(a3[0]); break;
  }
  std::memcpy(m16, t.m.m, 64);
  std::memcpy(minv16, t.m_inv.m, 64);
}
// Matrix4x4 / Transform algebra hooks for the reference's own unit tests
// (transform/matrix4x4.rs:208-353, transform/transform.rs:300-590).
void orc_mat_mul(const float* a16, const float* b16, float* out16) {
  M44 a, b;
  std::memcpy(a.m, a16, 64);
  std::memcpy(b.m, b16, 64);
  M44 r = a * b;
  std::memcpy(out16, r.m, 64);
}
void orc_mat_transpose(const float* a16, float* out16) {
  M44 a;
  std::memcpy(a.m, a16, 64);
  M44 r = a.transpose();
  std::memcpy(out16, r.m, 64);
}
void orc_xf_mul(const float* am, const float* aminv, const float* bm, const float* bminv, float* m16,
                float* minv16) {
  Transform t = xf_from(am, aminv) * xf_from(bm, bminv);
  std::memcpy(m16, t.m.m, 64);
  std::memcpy(minv16, t.m_inv.m, 64);
}
int orc_xf_swaps_handedness(const float* m, const float* minv) { return xf_from(m, minv).swaps_handedness() ? 1 : 0; }
void orc_xf_bbox(const float* m, const float* minv, const float* box6, float* out6) {
  BBox b = xf_bbox(xf_from(m, minv), BBox(V3(box6[0], box6[1], box6[2]), V3(box6[3], box6[4], box6[5])));
  out6[0] = b.p_min.x; out6[1] = b.p_min.y; out6[2] = b.p_min.z;
  out6[3] = b.p_max.x; out6[4] = b.p_max.y; out6[5] = b.p_max.z;
}
// Transform::xf(Ray): ray8 = o, mint, d, maxt
void orc_xf_ray(const float* m, const float* minv, const float* ray8, float* out8) {
  Ray r(V3(ray8[0], ray8[1], ray8[2]), V3(ray8[4], ray8[5], ray8[6]), ray8[3]);
  r.maxt = ray8[7];
  Ray t = xf_ray(xf_from(m, minv), r);
  out8[0] = t.o.x; out8[1] = t.o.y; out8[2] = t.o.z; out8[3] = t.mint;
  out8[4] = t.d.x; out8[5] = t.d.y; out8[6] = t.d.z; out8[7] = t.maxt;
}
// geometry/{vector,point,normal}.rs primitives used all over the path.  op: 0 cross -> out3,
// 1 dot -> out[0], 2 length(a) -> out[0], 3 normalize(a) -> out3, 4 coordinate_system(a) -> out6,
// 5 distance(a, b) -> out[0], 6 face_forward(a, b) -> out3, 7 length_squared(a), 8 distance_squared
void orc_vec_op(int op, const float* a3, const float* b3, float* out) {
  V3 a(a3[0], a3[1], a3[2]), b(b3[0], b3[1], b3[2]), r;
  switch (op) {
    case 0: r = cross(a, b); break;
    case 1: out[0] = dot(a, b); return;
    case 2: out[0] = length(a); return;
    case 3: r = normalize(a); break;
    case 4: {
      V3 x, y;
      coordinate_system(a, &x, &y);
      out[0] = x.x; out[1] = x.y; out[2] = x.z; out[3] = y.x; out[4] = y.y; out[5] = y.z;
      return;
    }
    case 5: out[0] = distance(a, b); return;
    case 6: r = face_forward(a, b); break;
    case 7: out[0] = length_squared(a); return;
    default: out[0] = length_squared(a - b); return;
  }
  out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void orc_task_windows(const int32_t* ext4, uint32_t num_tasks, int32_t* windows, uint32_t* keys) {
  SamplerDesc sd;
  for (int i = 0; i < 4; ++i) sd.ext[i] = ext4[i];
  auto tw = task_windows(sd, num_tasks);
  for (uint32_t t = 0; t < num_tasks; ++t) {
    for (int i = 0; i < 4; ++i) windows[4 * t + i] = tw[t].ext[i];
    for (int i = 0; i < 8; ++i) keys[8 * t + i] = tw[t].key[i];
  }
}
// write_img's pixel quantisation (camera/film.rs:21-23):
//   (255.0 * p.powf(1.0 / 2.2) + 0.5).clamp(0.0, 255.0) as u8      (`as u8` saturates, NaN -> 0)
void orc_rgb_to_bytes(const float* rgb, uint64_t n, uint8_t* out) {
  for (uint64_t i = 0; i < n; ++i) {
    float v = 255.0f * std::pow(rgb[i], 1.0f / 2.2f) + 0.5f;
    if (v < 0.0f) v = 0.0f;   // f32::clamp keeps NaN; the cast below maps it to 0
    if (v > 255.0f) v = 255.0f;
    out[i] = std::isnan(v) ? (uint8_t)0 : (uint8_t)v;
  }
}
// ---- MIPMap known-answer hooks (texture/mipmap.rs) ----
void* orc_mipmap_new(const float* rgb, uint32_t w, uint32_t h, int spectrum, int do_trilinear,
                     float max_aniso, int wrap, float scale, float gamma) {
  return new MIPMap(make_image_mipmap(rgb, w, h, spectrum != 0, do_trilinear != 0, max_aniso, wrap, scale, gamma));
}
void orc_mipmap_free(void* m) { delete static_cast<MIPMap*>(m); }
uint32_t orc_mipmap_levels(void* m) { return (uint32_t)static_cast<MIPMap*>(m)->levels(); }
void orc_mipmap_level_size(void* m, uint32_t level, uint32_t* wh) {
  const MipLevel& l = static_cast<MIPMap*>(m)->pyramid[level];
  wh[0] = (uint32_t)l.w;
  wh[1] = (uint32_t)l.h;
}
void orc_mipmap_level(void* m, uint32_t level, float* out_rgb) {
  const MipLevel& l = static_cast<MIPMap*>(m)->pyramid[level];
  for (size_t i = 0; i < l.px.size(); ++i)
    for (int c = 0; c < 3; ++c) out_rgb[3 * i + c] = l.px[i].c[c];
}
// n lookups: st6 = (s, t, dsdx, dtdx, dsdy, dtdy) per lookup
void orc_mipmap_lookup(void* m, const float* st6, uint64_t n, float* out_rgb) {
  const MIPMap* mm = static_cast<MIPMap*>(m);
  for (uint64_t i = 0; i < n; ++i) {
    const float* q = st6 + 6 * i;
    RGB r = mm->lookup(q[0], q[1], q[2], q[3], q[4], q[5]);
    for (int c = 0; c < 3; ++c) out_rgb[3 * i + c] = r.c[c];
  }
}
// ImageTexture::eval through a PlanarMapping2D::new() ((1,0,0),(0,1,0),0,0) at points p with
// dpdx/dpdy — the exact call the reference's imagemap.rs tests make.  pd9 = p, dpdx, dpdy.
void orc_image_texture_eval_planar(void* m, const float* pd9, uint64_t n, float* out_rgb) {
  const MIPMap* mm = static_cast<MIPMap*>(m);
  Mapping2D mp;
  mp.kind = 1;
  for (uint64_t i = 0; i < n; ++i) {
    DiffGeom dg;
    const float* q = pd9 + 9 * i;
    dg.p = V3(q[0], q[1], q[2]);
    dg.dpdx = V3(q[3], q[4], q[5]);
    dg.dpdy = V3(q[6], q[7], q[8]);
    float o[6];
    mp.map(dg, o);
    RGB r = mm->lookup(o[0], o[1], o[2], o[3], o[4], o[5]);
    for (int c = 0; c < 3; ++c) out_rgb[3 * i + c] = r.c[c];
  }
}
float orc_sinc_1d(float x, float tau) { return sinc_1d(x, tau); }
int32_t orc_modulo(int32_t a, int32_t b) { return modulo(a, b); }
// Texture::evaluate / TextureMapping2D::map at a hand-built DifferentialGeometry (the texture tests
// of texture/{mod,checkerboard,uv,mapping2d}.rs).  dg15 = p, dpdx, dpdy, u, v, dudx, dudy, dvdx, dvdy.
static DiffGeom dg_from15(const float* q) {
  DiffGeom dg;
  dg.p = V3(q[0], q[1], q[2]);
  dg.dpdx = V3(q[3], q[4], q[5]);
  dg.dpdy = V3(q[6], q[7], q[8]);
  dg.u = q[9];
  dg.v = q[10];
  dg.dudx = q[11];
  dg.dudy = q[12];
  dg.dvdx = q[13];
  dg.dvdy = q[14];
  return dg;
}
void orc_texture_eval(OrcScene* s, int tex_id, const float* dg15, float* out3) {
  RGB r = s->sc.textures.eval(tex_id, dg_from15(dg15));
  out3[0] = r.c[0];
  out3[1] = r.c[1];
  out3[2] = r.c[2];
}
void orc_mapping_map(int map_kind, const float* map12, const float* dg15, float* out6) {
  mapping_from(map_kind, map12).map(dg_from15(dg15), out6);
}
// IdentityMapping3D::map (mapping3d.rs:57-63): out9 = p, dpdx, dpdy
void orc_mapping3d_map(const float* map12, const float* dg15, float* out9) {
  V3 p, dx, dy;
  mapping_from(4, map12).map3(dg_from15(dg15), &p, &dx, &dy);
  out9[0] = p.x; out9[1] = p.y; out9[2] = p.z;
  out9[3] = dx.x; out9[4] = dx.y; out9[5] = dx.z;
  out9[6] = dy.x; out9[7] = dy.y; out9[8] = dy.z;
}
// texture/noise.rs: noise (:70-103), fbm (:112-127), turbulence (:129-145)
float orc_noise(float x, float y, float z) { return noise(x, y, z); }
float orc_fbm(int turbulence, const float* p3, const float* dpdx3, const float* dpdy3, float omega, int octaves) {
  return fbm_or_turbulence(turbulence != 0, V3(p3[0], p3[1], p3[2]), V3(dpdx3[0], dpdx3[1], dpdx3[2]),
                           V3(dpdy3[0], dpdy3[1], dpdy3[2]), omega, octaves);
}
// material::bump (material/mod.rs:23-77) at a hand-built shading geometry.
// dgs33 = p, dpdu, dpdv, dndu, dndv, nn, (u, v, dudx, dudy, dvdx, dvdy), (flip, 0, 0), dpdx, dpdy;
// ng3 = geometric normal.
// out9 = bumped dpdu, dpdv, nn.
void orc_bump(OrcScene* s, int tex_id, const float* dgs33, const float* ng3, float* out9) {
  DiffGeom g;
  const float* q = dgs33;
  g.p = V3(q[0], q[1], q[2]);
  g.dpdu = V3(q[3], q[4], q[5]);
  g.dpdv = V3(q[6], q[7], q[8]);
  g.dndu = V3(q[9], q[10], q[11]);
  g.dndv = V3(q[12], q[13], q[14]);
  g.nn = V3(q[15], q[16], q[17]);
  g.u = q[18]; g.v = q[19]; g.dudx = q[20]; g.dudy = q[21]; g.dvdx = q[22]; g.dvdy = q[23];
  g.flip = q[24] != 0.0f;
  g.dpdx = V3(q[27], q[28], q[29]);
  g.dpdy = V3(q[30], q[31], q[32]);
  DiffGeom gg;
  gg.nn = V3(ng3[0], ng3[1], ng3[2]);
  DiffGeom b = bump_dg(s->sc.textures, tex_id, gg, g);
  out9[0] = b.dpdu.x; out9[1] = b.dpdu.y; out9[2] = b.dpdu.z;
  out9[3] = b.dpdv.x; out9[4] = b.dpdv.y; out9[5] = b.dpdv.z;
  out9[6] = b.nn.x; out9[7] = b.nn.y; out9[8] = b.nn.z;
}
}  // extern "C"
