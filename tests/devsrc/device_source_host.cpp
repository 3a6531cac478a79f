// TEST INFRASTRUCTURE ONLY.  Compiles the device source of pbrt_rust_b200/csrc/ (the *_math.cuh /
// trace_core.cuh / shade*.cuh / film.cuh / raygen.cuh / halton.cuh headers, kernels included) as
// host code: PB_HOST_CHECK replaces the few CUDA intrinsics and launch keywords by plain C++ and
// runs a kernel as one emulated thread per block.  The CPU test-suite drives it against the oracle
// where no GPU exists — single functions, whole kernels and whole frames.  The product never loads
// this library.
#define PB_HOST_CHECK 1
#include "../../pbrt_rust_b200/csrc/shade_tex.cuh"
#include "../../pbrt_rust_b200/csrc/shade_mip.cuh"

#include <cstring>

static_assert(sizeof(DG) == 30 * sizeof(float), "DG is 30 packed floats");

extern "C" {
// dg30 = struct DG: p, nn, u, v, dpdu, dpdv, dndu, dndv, dpdx, dpdy, dudx, dudy, dvdx, dvdy
void devsrc_tex_eval_img(const pbrtb200_texture* table, const pbrtb200_mipmap* mipmaps, const float* texels, int id,
                         const float* dg30, float* out3) {
  DG dg;
  std::memcpy(&dg, dg30, sizeof dg);
  TexEnv env{table, mipmaps, reinterpret_cast<const float4*>(texels)};
  const f3 r = tex_eval_ext<PBRTB200_TEX_MAX_DEPTH>(env, id, dg);
  out3[0] = r.x;
  out3[1] = r.y;
  out3[2] = r.z;
}
void devsrc_tex_eval(const pbrtb200_texture* table, int id, const float* dg30, float* out3) {
  DG dg;
  std::memcpy(&dg, dg30, sizeof dg);
  TexEnv env{table, nullptr, nullptr};
  const f3 r = tex_eval_ext<PBRTB200_TEX_MAX_DEPTH>(env, id, dg);
  out3[0] = r.x;
  out3[1] = r.y;
  out3[2] = r.z;
}
void devsrc_tex_map(const pbrtb200_texture* tx, const float* dg30, float* out6) {
  DG dg;
  std::memcpy(&dg, dg30, sizeof dg);
  tex_map_ext(*tx, dg, out6);
}
// out9 = bumped dpdu, dpdv, nn
void devsrc_bump(const pbrtb200_texture* table, int id, const float* dg30, const float* ng3, int flip,
                 float* out9) {
  DG dg;
  std::memcpy(&dg, dg30, sizeof dg);
  TexEnv env{table, nullptr, nullptr};
  const DG b = bump_dg_(env, id, dg, mk3(ng3[0], ng3[1], ng3[2]), flip != 0);
  out9[0] = b.dpdu.x; out9[1] = b.dpdu.y; out9[2] = b.dpdu.z;
  out9[3] = b.dpdv.x; out9[4] = b.dpdv.y; out9[5] = b.dpdv.z;
  out9[6] = b.nn.x;   out9[7] = b.nn.y;   out9[8] = b.nn.z;
}
float devsrc_noise(float x, float y, float z) { return noise_(x, y, z); }
float devsrc_fbm(int turb, const float* p, const float* dpdx, const float* dpdy, float omega, int octaves) {
  return fbm_(turb != 0, mk3(p[0], p[1], p[2]), mk3(dpdx[0], dpdx[1], dpdx[2]), mk3(dpdy[0], dpdy[1], dpdy[2]),
              omega, octaves);
}
}

// ---- HaltonSampler arithmetic (pbrt_rust_b200/csrc/halton_math.cuh) -------------------------------
#include "../../pbrt_rust_b200/csrc/halton_math.cuh"
extern "C" {
double devsrc_radical_inverse(unsigned long long n, unsigned int b) { return radical_inverse_(n, b); }
// win4 = x0, x1, y0, y1 of the task window; returns 1 and the image position when the candidate is kept
int devsrc_halton_image(const int* win4, float delta, unsigned long long i, float* out2) {
  DHaltonTask t{};
  t.x0 = win4[0]; t.x1 = win4[1]; t.y0 = win4[2]; t.y1 = win4[3];
  t.delta = delta;
  return halton_image(t, i, &out2[0], &out2[1]) ? 1 : 0;
}
unsigned int devsrc_halton_prime(int k) { return pb_halton_primes[k]; }
}

// ---- traversal arithmetic (csrc/trace_math.cuh) and dmath.cuh helpers ------------------------------
#include "../../pbrt_rust_b200/csrc/trace_math.cuh"
extern "C" {
// box6 = bmin, bmax; ray8 = o, mint, d, maxt.  finite != 0: the min/max variant for finite 1/d.
int devsrc_slab(const float* box6, const float* ray8, int finite, float* t0_out) {
  const f3 o = mk3(ray8[0], ray8[1], ray8[2]);
  const f3 inv = mk3(1.f / ray8[4], 1.f / ray8[5], 1.f / ray8[6]);
  return (finite ? slab_test_finite(box6[0], box6[1], box6[2], box6[3], box6[4], box6[5], o, inv, ray8[3], ray8[7], t0_out)
                 : slab_test(box6[0], box6[1], box6[2], box6[3], box6[4], box6[5], o, inv, ray8[3], ray8[7], t0_out))
             ? 1 : 0;
}
int devsrc_tri_hit(const float* p9, const float* ray8, float* tbb) {
  return tri_hit(mk3(p9[0], p9[1], p9[2]), mk3(p9[3], p9[4], p9[5]), mk3(p9[6], p9[7], p9[8]), mk3(ray8[0], ray8[1], ray8[2]),
                 mk3(ray8[4], ray8[5], ray8[6]), ray8[3], ray8[7], &tbb[0], &tbb[1], &tbb[2]) ? 1 : 0;
}
int devsrc_quadratic(float a, float b, float c, float* t01) { return quadratic_(a, b, c, &t01[0], &t01[1]) ? 1 : 0; }
int devsrc_solve2x2(const float* a4, const float* b2, float* x2) {
  return solve2x2_(a4[0], a4[1], a4[2], a4[3], b2[0], b2[1], &x2[0], &x2[1]) ? 1 : 0;
}
void devsrc_coordinate_system(const float* v3, float* out6) {
  f3 a, b;
  coordinate_system_(mk3(v3[0], v3[1], v3[2]), &a, &b);
  out6[0] = a.x; out6[1] = a.y; out6[2] = a.z; out6[3] = b.x; out6[4] = b.y; out6[5] = b.z;
}
void devsrc_chacha12_block(const uint32_t* key8, unsigned long long blk, uint32_t* out16) { chacha12_block(key8, blk, out16); }
}

// ---- shading arithmetic (csrc/shade_math.cuh) ------------------------------------------------------
#include "../../pbrt_rust_b200/csrc/shade_math.cuh"
extern "C" {
// the material -> DBSDF step of k_shade (matte.rs:30-51, plastic.rs:32-55) for constant textures,
// then BSDF::f.  frame9 = shading nn, geometric ng, dpdu.
void devsrc_bsdf_f(int kind, const float* kd3, const float* ks3, float sigma_or_rough, const float* frame9,
                   const float* wo3, const float* wi3, int strict_flags, float* out3) {
  DBSDF bs;
  bs.nn = mk3(frame9[0], frame9[1], frame9[2]);
  bs.ng = mk3(frame9[3], frame9[4], frame9[5]);
  bs.tn = bs.sn = mk3(0.f, 0.f, 0.f);
  bs.kd = mk3(kd3[0], kd3[1], kd3[2]);
  bs.ks = mk3(0.f, 0.f, 0.f);
  bs.a = bs.b = 0.f;
  bs.kind = kind;
  if (kind == 1) {
    const float sigma = sigma_or_rough * PB_PI / 180.0f;
    const float sigma2 = sigma * sigma;
    bs.a = 1.0f - (sigma2 / (2.0f * (sigma + 0.33f)));
    bs.b = 0.45f * sigma2 / (sigma2 + 0.09f);
  } else if (kind == 2) {
    bs.ks = mk3(ks3[0], ks3[1], ks3[2]);
    float e = 1.0f / sigma_or_rough;
    if (e > 1000.0f || isnan(e)) e = 1000.0f;
    bs.a = e;
  }
  if (bs.kind != 0) {
    bs.tn = normalize3(mk3(frame9[6], frame9[7], frame9[8]));
    bs.sn = cross3(bs.nn, bs.tn);
  }
  const f3 f = bsdf_f(bs, mk3(wo3[0], wo3[1], wo3[2]), mk3(wi3[0], wi3[1], wi3[2]), strict_flags != 0);
  out3[0] = f.x; out3[1] = f.y; out3[2] = f.z;
}
float devsrc_fresnel_dielectric(float cosi, float ei, float et) { return fresnel_dielectric_(cosi, ei, et); }
void devsrc_compute_differentials(const float* g12, const float* rd12, float* out10) {
  DG g{};
  g.p = mk3(g12[0], g12[1], g12[2]);
  g.nn = mk3(g12[3], g12[4], g12[5]);
  g.dpdu = mk3(g12[6], g12[7], g12[8]);
  g.dpdv = mk3(g12[9], g12[10], g12[11]);
  dg_compute_differentials(g, mk3(rd12[0], rd12[1], rd12[2]), mk3(rd12[3], rd12[4], rd12[5]),
                           mk3(rd12[6], rd12[7], rd12[8]), mk3(rd12[9], rd12[10], rd12[11]));
  out10[0] = g.dpdx.x; out10[1] = g.dpdx.y; out10[2] = g.dpdx.z;
  out10[3] = g.dpdy.x; out10[4] = g.dpdy.y; out10[5] = g.dpdy.z;
  out10[6] = g.dudx; out10[7] = g.dvdx; out10[8] = g.dudy; out10[9] = g.dvdy;
}
void devsrc_vis_segment(const float* p1, float eps1, const float* p2, float eps2, float* ray8) {
  pbrtb200_ray32 r;
  vis_segment(mk3(p1[0], p1[1], p1[2]), eps1, mk3(p2[0], p2[1], p2[2]), eps2, &r);
  ray8[0] = r.o[0]; ray8[1] = r.o[1]; ray8[2] = r.o[2]; ray8[3] = r.mint;
  ray8[4] = r.d[0]; ray8[5] = r.d[1]; ray8[6] = r.d[2]; ray8[7] = r.maxt;
}
// dg of a quadric hit: rec = the flattened record, o2w12 = object-to-world rows; out14 = p, nn, u, v, dpdu, dpdv
void devsrc_quadric_dg(const pbrtb200_sphere80* rec, const float* o2w12, const float* o3, const float* d3, float t_hit,
                       float phi, float* out14) {
  const DG g = sphere_dg(*rec, o2w12, mk3(o3[0], o3[1], o3[2]), mk3(d3[0], d3[1], d3[2]), t_hit, phi);
  const float v[14] = {g.p.x, g.p.y, g.p.z, g.nn.x, g.nn.y, g.nn.z, g.u, g.v, g.dpdu.x, g.dpdu.y, g.dpdu.z,
                       g.dpdv.x, g.dpdv.y, g.dpdv.z};
  std::memcpy(out14, v, sizeof v);
}
}

// ---- sampler helpers of csrc/dmath.cuh: LD radical inverses, the StdRng word stream, RNG::shuffle ----
extern "C" {
float devsrc_van_der_corput(uint32_t n, uint32_t scramble) { return van_der_corput_(n, scramble); }
float devsrc_sobol2(uint32_t n, uint32_t scramble) { return sobol2_(n, scramble); }
// n floats of RNG::random_float() from the stream keyed by key8, starting at word `start`
void devsrc_stream_floats(const uint32_t* key8, unsigned long long start, unsigned long long n, float* out) {
  WordStream ws;
  ws.init(key8, start);
  for (unsigned long long i = 0; i < n; ++i) out[i] = ws.random_float();
}
// RNG::shuffle (rng.rs:23-33) of `count` groups of dims (1 or 2) floats
void devsrc_shuffle(const uint32_t* key8, unsigned long long start, float* v, uint32_t count, int dims) {
  WordStream ws;
  ws.init(key8, start);
  if (dims == 2) shuffle2(ws, reinterpret_cast<float2*>(v), count); else shuffle1(ws, v, count);
}
}

// ---- Film::add_sample arithmetic (csrc/film_math.cuh) ------------------------------------------------
#include "../../pbrt_rust_b200/csrc/film_math.cuh"
extern "C" {
// weight the sample at (sx, sy) adds to every film pixel (0 where it does not reach): out_w row-major
void devsrc_film_weights(const pbrtb200_film* film, float sx, float sy, float* out_w) {
  DFilm f{};
  f.x_start = film->x_pixel_start;
  f.y_start = film->y_pixel_start;
  f.x_count = film->x_pixel_count;
  f.y_count = film->y_pixel_count;
  f.xw = film->filter_xw;
  f.yw = film->filter_yw;
  f.inv_xw = 1.0f / film->filter_xw;  // filter.rs:12-19
  f.inv_yw = 1.0f / film->filter_yw;
  for (int y = 0; y < f.y_count; ++y)
    for (int x = 0; x < f.x_count; ++x) {
      int ti;
      out_w[(size_t)y * f.x_count + x] =
          film_sample_index(f, sx, sy, f.x_start + x, f.y_start + y, &ti) ? film->filter_table[ti] : 0.0f;
    }
}
}

// ---- perspective camera ray (camera_ray, csrc/trace_math.cuh) ----------------------------------------
extern "C" {
// cs5 = image_x, image_y, lens_u, lens_v, time -> ray6 = o, d (world space).  The DCamera is filled
// exactly as api.cu's fill_camera does.
void devsrc_camera_ray(const pbrtb200_camera* c, int spp, const float* cs5, float* ray6) {
  DCamera dc;
  std::memcpy(dc.r2c, c->raster_to_camera, 64);
  std::memcpy(dc.c2w, c->camera_to_world, 64);
  for (int i = 0; i < 3; ++i) {
    dc.dx[i] = c->dx_camera[i];
    dc.dy[i] = c->dy_camera[i];
  }
  dc.sopen = c->shutter_open;
  dc.sclose = c->shutter_close;
  dc.lens_radius = c->lens_radius;
  dc.focal_distance = c->focal_distance;
  dc.diff_scale = 1.0f / std::sqrt((float)spp);
  f3 o, d;
  camera_ray(dc, cs5[0], cs5[1], cs5[2], cs5[3], &o, &d, nullptr);
  ray6[0] = o.x; ray6[1] = o.y; ray6[2] = o.z; ray6[3] = d.x; ray6[4] = d.y; ray6[5] = d.z;
}
}

// ---- triangle dg + shading geometry (tri_dg, tri_shading_geometry, csrc/shade_math.cuh) -------------
extern "C" {
// pw9 = WORLD-space p1, p2, p3; mesh = the flattened per-mesh record; n9 / s9 / uv6 = the triangle's
// attribute record (object-space normals / tangents, uvs) or NULL.
void devsrc_tri_surface(const pbrtb200_mesh* mesh, const float* pw9, const float* n9, const float* s9, const float* uv6,
                        const float* o3, const float* d3, float t, float b1, float b2, float* out_dg14, float* out_dgs17) {
  TriData td;
  td.p1 = mk3(pw9[0], pw9[1], pw9[2]);
  td.p2 = mk3(pw9[3], pw9[4], pw9[5]);
  td.p3 = mk3(pw9[6], pw9[7], pw9[8]);
  td.mesh = 0;
  td.attr = 0;
  if (uv6) {
    for (int k = 0; k < 3; ++k) { td.uv[k][0] = uv6[2 * k]; td.uv[k][1] = uv6[2 * k + 1]; }
  } else {  // mesh.rs:74-87 default uvs
    td.uv[0][0] = 0.f; td.uv[0][1] = 0.f; td.uv[1][0] = 1.f; td.uv[1][1] = 0.f; td.uv[2][0] = 1.f; td.uv[2][1] = 1.f;
  }
  DScene sc{};
  sc.tri_n = n9;
  sc.tri_s = s9;
  const DG dg = tri_dg(td, mk3(o3[0], o3[1], o3[2]), mk3(d3[0], d3[1], d3[2]), t, b1, b2, mesh->flip != 0);
  const DG gs = tri_shading_geometry(sc, td, *mesh, dg);
  const float a[14] = {dg.p.x, dg.p.y, dg.p.z, dg.nn.x, dg.nn.y, dg.nn.z, dg.u, dg.v,
                       dg.dpdu.x, dg.dpdu.y, dg.dpdu.z, dg.dpdv.x, dg.dpdv.y, dg.dpdv.z};
  std::memcpy(out_dg14, a, sizeof a);
  const float b[17] = {gs.nn.x, gs.nn.y, gs.nn.z, gs.dpdu.x, gs.dpdu.y, gs.dpdu.z, gs.dpdv.x, gs.dpdv.y, gs.dpdv.z,
                       gs.dndu.x, gs.dndu.y, gs.dndu.z, gs.dndv.x, gs.dndv.y, gs.dndv.z, gs.u, gs.v};
  std::memcpy(out_dgs17, b, sizeof b);
}
}

// ---- the traversal itself (csrc/trace_core.cuh) over a flattened scene --------------------------------
#include <vector>

#include "../../pbrt_rust_b200/csrc/host_logic.hpp"
#include "../../pbrt_rust_b200/csrc/trace_core.cuh"
namespace {
int g_box = 3;  // box-test family of the traversal (trace_core.cuh child_box), devsrc_set_box
// The traversal part of pbrtb200_upload_scene (csrc/api.cu): pair nodes, leaf boxes, box flags, and
// the "a one-triangle leaf's box equals the bounds of its vertices" check that decides whether the
// leaf boxes must be read from memory (MULTI variants).
bool traversal_scene(const pbrtb200_scene* s, pbh::PairNodes* pn, DScene* sc, bool* sph, bool* multi) {
  if (pbh::build_pair_nodes(s->nodes, s->n_nodes, s->n_prims, pn)) return false;
  sc->nodes = reinterpret_cast<const float4*>(pn->pairs.data());
  sc->tris = reinterpret_cast<const float4*>(s->tris);
  *sph = s->n_spheres > 0 || s->leaf_prim != nullptr;
  *multi = pn->multi;
  if (!*multi && !*sph && s->tris)
    for (uint32_t i = 0; i < s->n_nodes && !*multi; ++i) {
      const pbrtb200_node32& nd = s->nodes[i];
      if (!nd.is_leaf) continue;
      const pbrtb200_tri48& t = s->tris[nd.offset];
      const float* v[3] = {t.p1, t.p2, t.p3};
      for (int a = 0; a < 3; ++a) {
        const float lo = std::fmin(std::fmin(v[0][a], v[1][a]), v[2][a]), hi = std::fmax(std::fmax(v[0][a], v[1][a]), v[2][a]);
        if (!(lo == nd.bmin[a]) || !(hi == nd.bmax[a])) *multi = true;
      }
    }
  sc->leaf_prim = *sph ? s->leaf_prim : nullptr;
  sc->leaf_count = pn->big_leaf ? pn->leaf_count.data() : nullptr;
  sc->leaf_boxes = (*multi || *sph) ? reinterpret_cast<const float4*>(pn->leaf_boxes.data()) : nullptr;
  sc->spheres = s->spheres;
  sc->sphere_o2w = s->sphere_o2w;
  sc->n_prims = s->n_prims;
  sc->root_ref = pn->root_ref;
  for (int i = 0; i < 3; ++i) {
    sc->root_bmin[i] = pn->root_bmin[i];
    sc->root_bmax[i] = pn->root_bmax[i];
    sc->babs[i] = pn->babs[i];
  }
  sc->boxes_finite = pn->boxes_finite ? 1u : 0u;
  sc->boxes_ordered = pn->boxes_ordered ? 1u : 0u;
  return true;
}
template <bool ANY, int MODE, int BOX>
TraceResult trace_variant_box(const DScene& sc, bool sph, bool multi, f3 o, f3 d, float mint, float maxt, uint32_t* sr, float* st) {
  return sph ? (multi ? trace_ray<ANY, true, true, MODE, BOX>(sc, o, d, mint, maxt, sr, st)
                      : trace_ray<ANY, true, false, MODE, BOX>(sc, o, d, mint, maxt, sr, st))
             : (multi ? trace_ray<ANY, false, true, MODE, BOX>(sc, o, d, mint, maxt, sr, st)
                      : trace_ray<ANY, false, false, MODE, BOX>(sc, o, d, mint, maxt, sr, st));
}
template <bool ANY, int MODE>
TraceResult trace_any_variant(const DScene& sc, bool sph, bool multi, f3 o, f3 d, float mint, float maxt, uint32_t* sr, float* st) {
  switch (g_box) {
    case 1: return trace_variant_box<ANY, MODE, 1>(sc, sph, multi, o, d, mint, maxt, sr, st);
    case 2: return trace_variant_box<ANY, MODE, 2>(sc, sph, multi, o, d, mint, maxt, sr, st);
    default: return trace_variant_box<ANY, MODE, 3>(sc, sph, multi, o, d, mint, maxt, sr, st);
  }
}
}  // namespace
extern "C" {
// Box-test family used by devsrc_trace / devsrc_render from now on (1, 2, 3; the product's
// PBRTB200_BOX).
void devsrc_set_box(int box) { g_box = box; }
// Bit a set: the specialised loops treat axis a as "the warp mixes signs here" (min / max form).
void devsrc_force_mixed_axes(int mask) { pb_host_force_mixed = mask; }
// Scene::intersect / intersect_p for n rays (ray8 = o, mint, d, maxt) through the device traversal source,
// with the pair nodes packed by the product's own build_pair_nodes and the kernel variant the library
// would launch.  any_mode: -1 closest hit (hit4 = prim, t, b1, b2 per ray), else the any-hit SIMT mode
// 0..4 (hit4[0] = prim or MISS).  Returns 0, or -1 with a bad scene / -2 on a stack overflow.
int devsrc_trace(const pbrtb200_scene* s, const float* rays8, unsigned long long n, int any_mode, float* hit4) {
  pbh::PairNodes pn;
  DScene sc{};
  bool sph = false, multi = false;
  if (!traversal_scene(s, &pn, &sc, &sph, &multi)) return -1;
  std::vector<uint32_t> s_ref((size_t)(PB_SM_STACK_MAX + 1) * PB_TRACE_THREADS);
  std::vector<float> s_t0((size_t)(PB_SM_STACK_MAX + 1) * PB_TRACE_THREADS);  // (host stack: two arrays)
  int rc = 0;
  for (unsigned long long i = 0; i < n; ++i) {
    const float* r = rays8 + 8 * i;
    const f3 o = mk3(r[0], r[1], r[2]), d = mk3(r[4], r[5], r[6]);
    uint32_t* sr = s_ref.data() + (i % PB_TRACE_THREADS);  // this "thread's" stack column
    float* st = s_t0.data() + (i % PB_TRACE_THREADS);
    TraceResult t;
#define PB_T(ANY, MODE) t = trace_any_variant<ANY, MODE>(sc, sph, multi, o, d, r[3], r[7], sr, st)
    switch (any_mode) {
      case -1: PB_T(false, 1); break;
      case 0: PB_T(true, 0); break;
      case 1: PB_T(true, 1); break;
      case 2: PB_T(true, 2); break;
      case 3: PB_T(true, 3); break;
      default: PB_T(true, 4); break;
    }
#undef PB_T
    if (t.prim == PB_OVERFLOW) rc = -2;
    hit4[4 * i] = __uint_as_float(t.prim);
    hit4[4 * i + 1] = t.t;
    hit4[4 * i + 2] = t.b1;
    hit4[4 * i + 3] = t.b2;
  }
  return rc;
}
}

// ---- the film gather itself (film_pixel, csrc/film.cuh) over a whole frame -----------------------------
#include "../../pbrt_rust_b200/csrc/film.cuh"
extern "C" {
// Whole-film gather of a frame whose sampler pixels are listed in raster order over the sampler extent
// ext4 = (x0, x1, y0, y1), spp samples each: img2 = image (x, y) per sample, rgb3 = its radiance (one
// non-area light term, no Le slot).  out_xyzw = the film, row-major over the film pixel extent.
void devsrc_film_gather(const pbrtb200_film* film, const int* ext4, int spp, const float* img2, const float* rgb3,
                        float* out_xyzw) {
  DFilm f{};
  f.x_start = film->x_pixel_start;
  f.y_start = film->y_pixel_start;
  f.x_count = film->x_pixel_count;
  f.y_count = film->y_pixel_count;
  f.xw = film->filter_xw;
  f.yw = film->filter_yw;
  f.inv_xw = 1.0f / film->filter_xw;
  f.inv_yw = 1.0f / film->filter_yw;
  f.sx0 = ext4[0]; f.sx1 = ext4[1]; f.sy0 = ext4[2]; f.sy1 = ext4[3];
  f.spp = spp;
  const size_t npx = (size_t)(ext4[1] - ext4[0]) * (size_t)(ext4[3] - ext4[2]), ns = npx * (size_t)spp;
  std::vector<float4> rec(ns);  // radiance records: one light term, no emitter
  for (size_t i = 0; i < ns; ++i) rec[i] = make_float4(rgb3[3 * i], rgb3[3 * i + 1], rgb3[3 * i + 2], 0.f);
  std::vector<uint32_t> edge(npx, 1u);
  std::vector<int32_t> index(npx);
  for (size_t i = 0; i < npx; ++i) index[i] = (int32_t)i;
  const int32_t rect[4] = {f.x_start, f.y_start, f.x_start + f.x_count, f.y_start + f.y_count};
  const uint32_t prefix[2] = {0u, (uint32_t)(f.x_count * f.y_count)};
  FilmArgs a{};
  a.img = reinterpret_cast<const float2*>(img2);
  a.rec = rec.data();
  a.offsets = nullptr;
  a.edge = edge.data();
  a.pix_index = index.data();
  a.rects = rect;
  a.rect_prefix = prefix;
  a.n_rects = 1;
  a.n_pixels = prefix[1];
  a.first = 0;
  a.count = prefix[1];
  a.pixel_mask = 0xFFFFFFFFu;
  a.out = reinterpret_cast<float4*>(out_xyzw);
  std::memcpy(a.table, film->filter_table, sizeof a.table);
  for (uint32_t gid = 0; gid < a.n_pixels; ++gid) film_pixel(f, a, gid);
}
}

// ---- a whole frame through the device source: camera rays, closest hit, k_shade, any-hit, film ----------
#include "../../pbrt_rust_b200/csrc/shade.cuh"
extern "C" {
// One frame.  The camera samples come from the caller (raster pixel order over the sampler extent
// ext4, spp per pixel): img2 image positions, lens2 lens positions or NULL, lightu = area_sample_pairs
// float pairs per sample or NULL.  Everything else is the device source: camera_ray, the traversal,
// the k_shade kernel (one emulated thread per block), the any-hit pass over its shadow queue and
// film_pixel; the scene set-up mirrors pbrtb200_upload_scene.  out_xyzw = the film; stats3 = camera
// hits, shadow rays, occluded shadow rays.  Returns 0 or a negative code.
int devsrc_render(const pbrtb200_scene* s, const pbrtb200_camera* c, const pbrtb200_film* film, const int* ext4, int spp,
                  const float* img2, const float* lens2, const float* lightu, int strict_flags, float* out_xyzw,
                  unsigned long long* stats3) {
  pbh::PairNodes pn;
  DScene sc{};
  bool sph = false, multi = false;
  if (!traversal_scene(s, &pn, &sc, &sph, &multi)) return -1;
  sc.meshes = s->meshes;
  sc.tri_uv = s->tri_uv;
  sc.tri_n = s->tri_n;
  sc.tri_s = s->tri_s;
  sc.materials = s->materials;
  sc.textures = s->textures;
  sc.mipmaps = s->n_mipmaps ? s->mipmaps : nullptr;
  sc.texels = s->n_mipmaps ? reinterpret_cast<const float4*>(s->texels) : nullptr;
  sc.n_lights = s->n_lights;
  // mat_flags, the EXT decision, light slots, area-light records: as pbrtb200_upload_scene does
  std::vector<uint8_t> mat_flags(std::max<uint32_t>(1u, s->n_materials), 0);
  bool ext = false;
  for (uint32_t i = 0; i < s->n_textures; ++i)
    if (s->textures[i].kind > PBRTB200_TEX_IMAGE ||
        (s->textures[i].kind != PBRTB200_TEX_CONSTANT && s->textures[i].map_kind > PBRTB200_MAP_PLANAR))
      ext = true;
  for (uint32_t i = 0; i < s->n_materials; ++i) {
    const pbrtb200_material& m = s->materials[i];
    auto varying = [&](int id) { return s->textures[id].kind != PBRTB200_TEX_CONSTANT; };
    bool v = varying(m.kd);
    if (m.kind == PBRTB200_MAT_MATTE) v = v || varying(m.sigma);
    if (m.kind == PBRTB200_MAT_PLASTIC) v = v || varying(m.ks) || varying(m.roughness);
    if (m.bump != 0) v = ext = true;
    mat_flags[i] = v ? 1 : 0;
  }
  sc.mat_flags = mat_flags.data();
  std::vector<pbrtb200_light> lights(s->lights, s->lights + s->n_lights);
  for (const pbrtb200_light& l : lights) {
    if (l.kind == PBRTB200_LIGHT_AREA) {
      sc.light_slots += (uint32_t)l.num_samples;
      sc.area_sample_pairs += (uint32_t)l.num_samples;
    } else {
      sc.light_slots += 1;
    }
  }
  std::vector<DAreaTri> at(std::max<uint32_t>(1u, s->n_area_prims));
  for (uint32_t i = 0; i < s->n_area_prims; ++i) {  // k_area_tri_setup, one emulated thread per block
    blockIdx.x = i;
    k_area_tri_setup(sc, s->area_prims, s->n_area_prims, at.data());
  }
  for (pbrtb200_light& l : lights) {
    if (l.kind != PBRTB200_LIGHT_AREA) continue;
    float total = 0.f;
    for (uint32_t k = 0; k < l.n_tris; ++k) total += at[l.first_tri + k].area;
    l.total_area = total;
    float acc = 0.f, lo = 0.f;
    for (uint32_t k = 0; k < l.n_tris; ++k) {
      acc += at[l.first_tri + k].area;
      float hi = acc / total;
      if (k + 1 == l.n_tris) hi = 1.0f;
      at[l.first_tri + k].cdf_lo = lo;
      at[l.first_tri + k].cdf_hi = hi;
      lo = hi;
    }
  }
  sc.lights = lights.data();
  DCamera dc;
  std::memcpy(dc.r2c, c->raster_to_camera, 64);
  std::memcpy(dc.c2w, c->camera_to_world, 64);
  for (int i = 0; i < 3; ++i) {
    dc.dx[i] = c->dx_camera[i];
    dc.dy[i] = c->dy_camera[i];
  }
  dc.sopen = c->shutter_open;
  dc.sclose = c->shutter_close;
  dc.lens_radius = c->lens_radius;
  dc.focal_distance = c->focal_distance;
  dc.diff_scale = 1.0f / std::sqrt((float)spp);

  const size_t npx = (size_t)(ext4[1] - ext4[0]) * (size_t)(ext4[3] - ext4[2]), n = npx * (size_t)spp;
  const float2* img = reinterpret_cast<const float2*>(img2);
  const float2* lens = (lens2 && c->lens_radius > 0.0f) ? reinterpret_cast<const float2*>(lens2) : nullptr;
  std::vector<uint32_t> s_ref((size_t)(PB_SM_STACK_MAX + 1) * PB_TRACE_THREADS);
  std::vector<float> s_t0((size_t)(PB_SM_STACK_MAX + 1) * PB_TRACE_THREADS);  // (host stack: two arrays)
  // closest hit per camera sample (the SRC = 1 path of k_trace: camera_ray, mint 0, maxt f32::MAX)
  std::vector<pbrtb200_hit16> hits(std::max<size_t>(1, n));
  int rc = 0;
  for (size_t i = 0; i < n; ++i) {
    f3 o, d;
    camera_ray(dc, img[i].x, img[i].y, lens ? lens[i].x : 0.f, lens ? lens[i].y : 0.f, &o, &d, nullptr);
    const TraceResult t = trace_any_variant<false, 1>(sc, sph, multi, o, d, 0.0f, PB_F32_MAX, s_ref.data(), s_t0.data());
    if (t.prim == PB_OVERFLOW) rc = -2;
    hits[i].prim = t.prim;
    hits[i].t = t.t;
    hits[i].b1 = t.b1;
    hits[i].b2 = t.b2;
  }
  // k_shade: `slots` radiance terms per sample (one slot: they are the radiance records themselves)
  const uint32_t slots = std::max(1u, sc.light_slots);
  std::vector<float4> terms(std::max<size_t>(1, n * slots));
  std::vector<float4> rec(std::max<size_t>(1, n));
  std::vector<pbrtb200_ray32> sq_rays(std::max<size_t>(1, n * slots));
  std::vector<uint32_t> sq_slots(std::max<size_t>(1, n * slots));
  uint32_t sq_count = 0, nan_count = 0;
  unsigned long long hit_total = 0, occluded = 0;
  if (sc.n_lights) {
    ShadeArgs sa{};
    sa.img = img;
    sa.lens = lens;
    sa.lightu = sc.area_sample_pairs ? reinterpret_cast<const float2*>(lightu) : nullptr;
    sa.hits = hits.data();
    sa.area_tris = at.data();
    sa.terms = terms.data();
    sa.sq_rays = sq_rays.data();
    sa.sq_slots = sq_slots.data();
    sa.sq_count = &sq_count;
    sa.hit_total = &hit_total;
    sa.nan_count = &nan_count;
    sa.n = n;
    sa.slots = slots;
    sa.strict_flags = strict_flags;
    for (size_t i = 0; i < n; ++i) {
      blockIdx.x = (unsigned)i;
      if (ext)
        k_shade<PB_SHADE_EXT_MIN_BLOCKS, true, true>(sc, dc, sa);
      else if (sc.mipmaps)
        k_shade<PB_SHADE_MIN_BLOCKS, true>(sc, dc, sa);
      else
        k_shade<PB_SHADE_MIN_BLOCKS, false>(sc, dc, sa);
    }
    // the any-hit pass over the shadow queue (k_trace<ANY>, default loop shape 2)
    for (uint32_t q = 0; q < sq_count; ++q) {
      const pbrtb200_ray32& r = sq_rays[q];
      const TraceResult t = trace_any_variant<true, 2>(sc, sph, multi, mk3(r.o[0], r.o[1], r.o[2]), mk3(r.d[0], r.d[1], r.d[2]),
                                                       r.mint, r.maxt, s_ref.data(), s_t0.data());
      if (t.prim == PB_OVERFLOW) rc = -2;
      float4& term = terms[sq_slots[q] & PB_SQ_INDEX];  // what k_trace<ANY> does with the guarded term
      if (t.prim != PBRTB200_MISS) {
        term = make_float4(0.f, 0.f, 0.f, (sq_slots[q] & PB_SQ_KEEPW) ? term.w : 0.f);
        ++occluded;
      } else if (sq_slots[q] & PB_SQ_NAN) {
        ++nan_count;
      }
    }
  }
  DFold fd{};
  fd.slots = slots;
  fd.n_lights = sc.n_lights;
  for (uint32_t i = 0; i < sc.n_lights; ++i) {
    fd.area[i] = lights[i].kind == PBRTB200_LIGHT_AREA ? 1 : 0;
    fd.ns[i] = (uint16_t)(fd.area[i] ? lights[i].num_samples : 1);
  }
  if (sc.n_lights) {
    if (slots > 1) {  // k_fold
      for (size_t i = 0; i < n; ++i)
        if (fold_terms(fd, sc.lights, terms.data() + i * slots, &rec[i])) ++nan_count;
    } else {
      rec = terms;
    }
  }
  // film
  DFilm f{};
  f.x_start = film->x_pixel_start;
  f.y_start = film->y_pixel_start;
  f.x_count = film->x_pixel_count;
  f.y_count = film->y_pixel_count;
  f.xw = film->filter_xw;
  f.yw = film->filter_yw;
  f.inv_xw = 1.0f / film->filter_xw;
  f.inv_yw = 1.0f / film->filter_yw;
  f.sx0 = ext4[0]; f.sx1 = ext4[1]; f.sy0 = ext4[2]; f.sy1 = ext4[3];
  f.spp = spp;
  std::vector<uint32_t> edge(npx, 1u);
  std::vector<int32_t> index(npx);
  for (size_t i = 0; i < npx; ++i) index[i] = (int32_t)i;
  const int32_t rect[4] = {f.x_start, f.y_start, f.x_start + f.x_count, f.y_start + f.y_count};
  const uint32_t prefix[2] = {0u, (uint32_t)(f.x_count * f.y_count)};
  FilmArgs a{};
  a.img = img;
  a.rec = sc.n_lights ? rec.data() : nullptr;
  a.lights = sc.lights;
  a.edge = edge.data();
  a.pix_index = index.data();
  a.rects = rect;
  a.rect_prefix = prefix;
  a.n_rects = 1;
  a.n_pixels = prefix[1];
  a.first = 0;
  a.count = prefix[1];
  a.pixel_mask = 0xFFFFFFFFu;
  a.out = reinterpret_cast<float4*>(out_xyzw);
  std::memcpy(a.table, film->filter_table, sizeof a.table);
  for (uint32_t gid = 0; gid < a.n_pixels; ++gid) film_pixel(f, a, gid);
  stats3[0] = hit_total;
  stats3[1] = sq_count;
  stats3[2] = occluded;
  return nan_count ? -3 : rc;
}
}

// ---- the sample-generation kernels (k_raygen_groups / k_raygen_full, k_halton_*), emulated thread by thread ----
#include "../../pbrt_rust_b200/csrc/raygen.cuh"
#include "../../pbrt_rust_b200/csrc/halton.cuh"
namespace {
// the raster-order pixel list of a whole sampler extent with each pixel's task and in-window index
// (what build_pixel_list derives from sampler/base.rs:29-48), plus the task keys
void raster_pixel_list(const pbrtb200_sampler* smp, std::vector<DPixel>* list, std::vector<uint32_t>* keys,
                       std::vector<DHaltonTask>* htasks) {
  const int32_t ext[4] = {smp->x_start, smp->x_end, smp->y_start, smp->y_end};
  const int sw = ext[1] - ext[0], sh = ext[3] - ext[2];
  list->assign((size_t)sw * sh, DPixel{0, 0, 0xFFFFu});
  keys->assign(8 * (size_t)smp->num_tasks, 0u);
  unsigned long long first = 0;
  for (int t = 0; t < smp->num_tasks; ++t) {
    pbh::task_key((uint64_t)t, &(*keys)[8 * (size_t)t]);
    int32_t w[4];
    pbh::sampler_sub_window(ext, (uint64_t)t, (uint64_t)smp->num_tasks, w);
    DHaltonTask ht{};
    ht.x0 = w[0]; ht.x1 = w[1]; ht.y0 = w[2]; ht.y1 = w[3];
    ht.first = first;
    if (w[0] != w[1] && w[2] != w[3]) {
      const int dx = w[1] - w[0], dy = w[3] - w[2], m = dx > dy ? dx : dy;
      ht.wanted = (unsigned long long)((long long)m * (long long)m) * (unsigned long long)smp->xs;
      ht.delta = std::fmax((float)dy, (float)dx);
      const uint32_t tw = (uint32_t)(w[1] - w[0]);
      for (int y = w[2]; y < w[3]; ++y)
        for (int x = w[0]; x < w[1]; ++x) {
          DPixel& p = (*list)[(size_t)(y - ext[2]) * sw + (size_t)(x - ext[0])];
          p.xy = (int32_t)(((uint32_t)(uint16_t)(int16_t)x) | ((uint32_t)(uint16_t)(int16_t)y << 16));
          p.k = (uint32_t)(y - w[2]) * tw + (uint32_t)(x - w[0]);
          p.task = (uint32_t)t;
        }
    }
    first += ht.wanted;
    htasks->push_back(ht);
  }
}
}  // namespace
extern "C" {
// Stratified / LD sample generation for the whole sampler extent (raster pixel order, spp per pixel).
// full != 0 runs k_raygen_full (lens / time shuffles, LD), else k_raygen_groups (stratified, image
// samples + light floats only).  out_cs5: image_x, image_y, lens_u, lens_v, time; out_lightu: light_pairs
// float pairs per sample; out_edge: per pixel.  Film geometry for the edge flags comes from `film`.
void devsrc_raygen(const pbrtb200_sampler* smp, int light_pairs, const pbrtb200_film* film, int full, float* out_cs5,
                   float* out_lightu, uint32_t* out_edge) {
  std::vector<DPixel> list;
  std::vector<uint32_t> keys;
  std::vector<DHaltonTask> unused;
  raster_pixel_list(smp, &list, &keys, &unused);
  DSampler ds{};  // as fill_sampler does
  ds.kind = smp->kind;
  ds.xs = smp->xs;
  ds.ys = smp->kind == PBRTB200_SAMPLER_STRATIFIED ? smp->ys : 1;
  ds.jitter = smp->jitter;
  int spp = smp->xs * smp->ys;
  if (smp->kind != PBRTB200_SAMPLER_STRATIFIED) {
    spp = 1;
    while (spp < smp->xs) spp <<= 1;
  }
  ds.spp = spp;
  const uint32_t n = (uint32_t)spp;
  ds.cam_words = smp->kind == PBRTB200_SAMPLER_STRATIFIED ? (smp->jitter ? 9u * n : 4u * n) : 2u * (5u + 6u * n);
  ds.words_per_pixel = ds.cam_words + 2u * n * (uint32_t)light_pairs;
  ds.sopen = smp->shutter_open;
  ds.sclose = smp->shutter_close;
  ds.task_keys = keys.data();
  const size_t npx = list.size(), ns = npx * (size_t)spp;
  std::vector<float2> img(ns), lens(ns, float2{0.f, 0.f}), lu(std::max<size_t>(1, ns * (size_t)light_pairs));
  std::vector<float> tm(ns, 0.f);
  std::vector<uint32_t> edge(npx, 0u);
  RaygenArgs ra{};
  ra.pixels = list.data();
  ra.n_pixels = npx;
  ra.img = img.data();
  ra.lens = full ? lens.data() : nullptr;
  ra.time = full ? tm.data() : nullptr;
  ra.light_pairs = (uint32_t)light_pairs;
  ra.lightu = light_pairs ? lu.data() : nullptr;
  ra.edge = edge.data();
  ra.fx_start = film->x_pixel_start;
  ra.fy_start = film->y_pixel_start;
  ra.fx_count = film->x_pixel_count;
  ra.fy_count = film->y_pixel_count;
  ra.xw = film->filter_xw;
  ra.yw = film->filter_yw;
  const size_t threads = full ? npx : npx * (size_t)((spp + 7) / 8);
  for (size_t t = 0; t < threads; ++t) {
    blockIdx.x = (unsigned)t;
    if (full) k_raygen_full(ds, ra); else k_raygen_groups(ds, ra);
  }
  for (size_t i = 0; i < ns; ++i) {
    out_cs5[5 * i] = img[i].x;
    out_cs5[5 * i + 1] = img[i].y;
    out_cs5[5 * i + 2] = lens[i].x;
    out_cs5[5 * i + 3] = lens[i].y;
    out_cs5[5 * i + 4] = tm[i];
  }
  if (light_pairs) std::memcpy(out_lightu, lu.data(), ns * (size_t)light_pairs * sizeof(float2));
  std::memcpy(out_edge, edge.data(), npx * sizeof(uint32_t));
}

// The Halton kernels over the whole sampler extent: k_halton_bin<0>, the host scan, k_halton_bin<1>,
// k_halton_samples.  out_counts per pixel (raster); samples written compactly in pixel order: out_cs5
// (sum of counts entries), out_lightu (light_pairs float pairs each).  Returns the number of samples.
unsigned long long devsrc_halton(const pbrtb200_sampler* smp, int light_pairs, uint32_t* out_counts, float* out_cs5,
                                 float* out_lightu, unsigned long long capacity) {
  std::vector<DPixel> list;
  std::vector<uint32_t> keys;
  std::vector<DHaltonTask> tasks;
  raster_pixel_list(smp, &list, &keys, &tasks);
  const size_t npx = list.size();
  std::vector<int32_t> index(npx);
  for (size_t i = 0; i < npx; ++i) index[i] = (int32_t)i;
  std::vector<uint32_t> counts(npx, 0u), fill(npx, 0u), offsets(npx + 1, 0u);
  HaltonArgs a{};
  a.tasks = tasks.data();
  a.n_tasks = (uint32_t)tasks.size();
  a.n_candidates = tasks.empty() ? 0 : tasks.back().first + tasks.back().wanted;
  a.pix_index = index.data();
  a.sx0 = smp->x_start;
  a.sy0 = smp->y_start;
  a.sw = smp->x_end - smp->x_start;
  a.counts = counts.data();
  a.fill = fill.data();
  a.offsets = offsets.data();
  for (unsigned long long g = 0; g < a.n_candidates; ++g) {
    blockIdx.x = (unsigned)g;
    k_halton_bin<0>(a);
  }
  unsigned long long total = 0;
  for (size_t i = 0; i < npx; ++i) {
    offsets[i] = (uint32_t)total;
    total += counts[i];
  }
  offsets[npx] = (uint32_t)total;
  std::memcpy(out_counts, counts.data(), npx * sizeof(uint32_t));
  if (total > capacity) return total;
  std::vector<uint32_t> idx(std::max<unsigned long long>(1, total));
  a.idx = idx.data();
  for (unsigned long long g = a.n_candidates; g-- > 0;) {  // reverse order: the per-pixel sort must restore index order
    blockIdx.x = (unsigned)g;
    k_halton_bin<1>(a);
  }
  std::vector<float2> img(std::max<unsigned long long>(1, total)), lens(img.size()), lu(std::max<size_t>(1, img.size() * (size_t)light_pairs));
  std::vector<float> tm(img.size());
  HaltonSampleArgs sa{};
  sa.tasks = tasks.data();
  sa.pixels = list.data();
  sa.n_pixels = npx;
  sa.offsets = offsets.data();
  sa.idx = idx.data();
  sa.img = img.data();
  sa.lens = lens.data();
  sa.time = tm.data();
  sa.lightu = light_pairs ? lu.data() : nullptr;
  sa.light_pairs = (uint32_t)light_pairs;
  sa.edge = nullptr;
  sa.sopen = smp->shutter_open;
  sa.sclose = smp->shutter_close;
  for (size_t li = 0; li < npx; ++li) {
    blockIdx.x = (unsigned)li;
    k_halton_samples(sa);
  }
  for (unsigned long long i = 0; i < total; ++i) {
    out_cs5[5 * i] = img[i].x;
    out_cs5[5 * i + 1] = img[i].y;
    out_cs5[5 * i + 2] = lens[i].x;
    out_cs5[5 * i + 3] = lens[i].y;
    out_cs5[5 * i + 4] = tm[i];
  }
  if (light_pairs) std::memcpy(out_lightu, lu.data(), total * (size_t)light_pairs * sizeof(float2));
  return total;
}
}
