// Surface interaction, material/texture evaluation, light sampling and the Whitted light loop
// (sm_100a), one thread per camera sample that hit something.
//
// Replaces WhittedIntegrator::li (src/integrator/whitted.rs:30-66), Intersection::get_bsdf
// (src/intersection.rs:40-47), DifferentialGeometry::{new_with, compute_differentials}
// (src/diff_geom.rs:52-152), Triangle::intersect's dg part and get_shading_geometry
// (src/shape/mesh.rs:105-193, 220-262), Sphere::intersect's dg part (src/shape/sphere.rs:143-180)
// with compute_dg (src/shape/helpers.rs:17-43), MatteMaterial / PlasticMaterial::get_bsdf
// (src/material/matte.rs:30-51, plastic.rs:32-55), BSDF::{new_with_eta, world_to_local, f}
// (src/bsdf/mod.rs:70-149), Lambertian / OrenNayar / Microfacet(Blinn) / Fresnel::Dielectric
// (src/bsdf/{lambertian,orennayar,microfacet,fresnel}.rs), Constant / Checkerboard / UV textures
// with UVMapping2D / PlanarMapping2D (src/texture/*), PointLight / SpotLight::sample_l
// (src/light/{point,spot}.rs) and VisibilityTester::segment (src/visibility_tester.rs:16-24).
// Deviations from the as-written reference are exactly those of SURVEY §0.2 (D2, D3, D7, D8, D9,
// D10, D11) and are listed in DESIGN.md.
#pragma once
#include "scene.cuh"
#include "trace.cuh"
#include "shade_tex.cuh"  // DG, texture mappings, procedural textures, bump
#include "shade_mip.cuh"   // MIPMap::lookup (trilinear / EWA)
#include "shade_math.cuh"  // dg construction + differentials, quadric dg, BxDFs, BSDF::f, shadow segments

PB_DEV TriData load_tri(const DScene& sc, uint32_t tri) {
  TriData t;
  const float4* tp = sc.tris + 3ull * tri;
  const float4 a = ldg4(tp), b = ldg4(tp + 1), c = ldg4(tp + 2);
  t.p1 = mk3(a.x, a.y, a.z);
  t.p2 = mk3(b.x, b.y, b.z);
  t.p3 = mk3(c.x, c.y, c.z);
  t.mesh = __float_as_uint(a.w);
  t.attr = __float_as_uint(b.w);
  if (sc.tri_uv && sc.meshes[t.mesh].has_uv) {  // mesh.rs:74-87
    const float* q = sc.tri_uv + 6ull * t.attr;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      t.uv[k][0] = __ldg(q + 2 * k);
      t.uv[k][1] = __ldg(q + 2 * k + 1);
    }
  } else {
    t.uv[0][0] = 0.f; t.uv[0][1] = 0.f;
    t.uv[1][0] = 1.f; t.uv[1][1] = 0.f;
    t.uv[2][0] = 1.f; t.uv[2][1] = 1.f;
  }
  return t;
}

// ---- textures ---------------------------------------------------------------------------------
// Nested checkerboards are supported to depth 3 (validated at upload).
template <int DEPTH, bool IMG>
PB_DEV f3 tex_eval(const DScene& sc, int id, const DG& dg) {
  const pbrtb200_texture tx = sc.textures[id];
  if (tx.kind == PBRTB200_TEX_CONSTANT) return mk3(tx.value[0], tx.value[1], tx.value[2]);
  float m[6];
  tex_map(tx, dg, m);
  if (tx.kind == PBRTB200_TEX_UV)  // uv.rs:20-26
    return mk3(m[0] - floorf(m[0]), m[1] - floorf(m[1]), 0.0f);
  if (IMG && tx.kind == PBRTB200_TEX_IMAGE)  // imagemap.rs:200-206
    return mip_lookup(sc.texels, sc.mipmaps + tx.tex1, m[0], m[1], m[2], m[3], m[4], m[5]);
  if constexpr (DEPTH == 0) {
    return mk3(0.f, 0.f, 0.f);
  } else {
    const float s = m[0], t = m[1];
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
    if (point) return tex_eval<DEPTH - 1, IMG>(sc, first ? tx.tex1 : tx.tex2, dg);
    const float sint = ds > 0.0f ? (bump_int_(s1) - bump_int_(s0)) / (2.0f * ds) : 0.0f;
    const float tint = dt > 0.0f ? (bump_int_(t1) - bump_int_(t0)) / (2.0f * dt) : 0.0f;
    const float area_sq = (ds > 1.0f || dt > 1.0f) ? 0.5f : sint + tint - 2.0f * sint * tint;
    const f3 a = tex_eval<DEPTH - 1, IMG>(sc, tx.tex1, dg), b = tex_eval<DEPTH - 1, IMG>(sc, tx.tex2, dg);
    return a * (1.0f - area_sq) + b * area_sq;  // Lerp::lerp_with, utils/mod.rs:20
  }
}


// Warp-aggregated queue push: one atomicAdd per warp for all lanes that call it together.
PB_DEV uint32_t warp_agg_inc(uint32_t* ctr) {
  const unsigned mask = __activemask();
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(mask) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(ctr, (uint32_t)__popc(mask));
  base = __shfl_sync(mask, base, leader);
  return base + (uint32_t)__popc(mask & ((1u << lane) - 1u));
}

#ifndef PB_SHADE_MIN_BLOCKS
#define PB_SHADE_MIN_BLOCKS 8  // 64 registers: the stage is latency-bound, occupancy wins (profiles/r01_notes.md)
#endif

// Texture lookup of the shading kernel: the inlined evaluator of the common kinds, or the general
// out-of-line one (EXT).
template <bool IMG, bool EXT>
PB_DEV f3 shade_tex(const DScene& sc, const TexEnv& env, int id, const DG& dg) {
  if constexpr (EXT) {
    const pbrtb200_texture* tx = env.textures + id;  // constants (most sigma / ks / roughness maps) stay inline
    if (__ldg(&tx->kind) == PBRTB200_TEX_CONSTANT) return mk3(__ldg(&tx->value[0]), __ldg(&tx->value[1]), __ldg(&tx->value[2]));
    // The out-of-line evaluator takes the geometry by address; a private copy made only on this
    // path keeps the caller's `dg` in registers (an unconditional copy cost config 3 0.2 ms/frame
    // of local-memory stores although all of its textures are constants).
    const DG tmp = dg;
    return tex_eval_ext<PBRTB200_TEX_MAX_DEPTH>(env, id, tmp);
  } else {
    return tex_eval<3, IMG>(sc, id, dg);
  }
}

#ifndef PB_SHADE_EXT_MIN_BLOCKS
#define PB_SHADE_EXT_MIN_BLOCKS 6  // general texture evaluator + bump: 80 registers (4 / 6 / 8 measured, profiles/r01_notes.md)
#endif

struct ShadeArgs {
  const float2* __restrict__ img;
  const float2* __restrict__ lens;    // may be NULL
  const float2* __restrict__ lightu;  // light-sample float pairs (raygen), NULL without area lights
  const pbrtb200_hit16* __restrict__ hits;
  const DAreaTri* __restrict__ area_tris;
  // Per camera sample of the chunk `slots` float4 radiance terms (index idx * slots + j):
  //   term j = (f * Li * |wi.n| / pdf of light slot j, 0 when nothing is reflected ; w)
  //   w of term 0 = bits(e): e - 1 = the area light whose emitter was hit facing the camera, 0 = none
  // The any-hit kernel zeroes the xyz of a term whose shadow ray is occluded.  With ONE slot the terms
  // ARE the frame's radiance records (film.cuh: L = Le(e) + v); with several, k_fold folds them.
  float4* __restrict__ terms;
  pbrtb200_ray32* __restrict__ sq_rays;  // shadow-ray queue (chunk-local)
  uint32_t* __restrict__ sq_slots;       // term each shadow ray guards (chunk-local index) | PB_SQ_* flags
  uint32_t* sq_count;
  unsigned long long* hit_total;
  uint32_t* nan_count;                   // NaN emitted radiance (the terms are checked where they become final)
  uint64_t n;
  uint32_t slots;
  int strict_flags;
};

// IMG: the scene has image textures (a separate instantiation keeps the MIPMap code, its
// registers and its call-site spills out of the kernel every other scene runs).
// EXT: the scene uses spherical / cylindrical mappings, scale / mix / bilerp / dots / fbm /
// wrinkled textures or bump maps: every texture goes through the out-of-line general evaluator
// (shade_tex.cuh) and material::bump runs before the BSDF frame is built.
template <int MIN_BLOCKS, bool IMG, bool EXT = false>
__global__ void __launch_bounds__(128, MIN_BLOCKS)
k_shade(const DScene sc, const DCamera cam, const ShadeArgs a) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = idx < a.n;
  float4 hraw = make_float4(__uint_as_float(PBRTB200_MISS), 0.f, 0.f, 0.f);
  if (in_range) {
    hraw = ld_stream(reinterpret_cast<const float4*>(a.hits) + idx);
    // independent streams this thread will need later: start them now (no register cost)
    PB_PREFETCH_L2(a.img + idx);
    if (a.lightu) PB_PREFETCH_L2(a.lightu + idx * sc.area_sample_pairs);
  }
  const uint32_t prim = __float_as_uint(hraw.x);
  // Block-level bookkeeping. A same-address global atomic retires at ~0.66 ns on B200 (measured,
  // scripts/micro/atomic_bench.cu), so one atomic per WARP for the hit count and one per warp and
  // light slot for the shadow queue (2 M per config-3 frame) cost more than the shading arithmetic.
  // Both are aggregated per block in shared memory instead: one global atomic per block.
  __shared__ uint32_t s_hit_cnt[4];
  __shared__ uint32_t s_warp_cnt[2][4];  // double-buffered by slot parity: two barriers per slot suffice
  __shared__ uint32_t s_block_base[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  {
    const unsigned hm = __ballot_sync(0xffffffffu, prim != PBRTB200_MISS);
    if (lane == 0) s_hit_cnt[warp] = __popc(hm);
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t t = s_hit_cnt[0] + s_hit_cnt[1] + s_hit_cnt[2] + s_hit_cnt[3];
      if (t) atomicAdd(a.hit_total, (unsigned long long)t);
    }
  }
  // Every thread of the block walks the light loop below (its trip counts depend on the scene
  // only) so the queue push can use block barriers; `alive` threads are the ones with a hit.
  const bool alive = in_range && prim != PBRTB200_MISS;
  float4* terms = a.terms + (in_range ? idx : 0) * a.slots;
  if (in_range && !alive)  // miss: sum of light.le(ray) = 0 (light/mod.rs:50-52)
    for (uint32_t q = 0; q < a.slots; ++q) st_stream(terms + q, make_float4(0.f, 0.f, 0.f, 0.f));
  DBSDF bs;
  uint32_t le_bits = 0u;
  f3 p = mk3(0.f, 0.f, 0.f), n = p, wo = p;
  float ray_epsilon = 0.f;
  if (alive) {
  const float t_hit = hraw.y, hb1 = hraw.z, hb2 = hraw.w;
  // The triangle record is the second link of a dependent load chain (hit -> triangle -> mesh ->
  // material -> texture); start fetching it now so it overlaps the camera-ray arithmetic below.
  if (!sc.leaf_prim) {
    const char* tp = reinterpret_cast<const char*>(sc.tris + 3ull * prim);
    PB_PREFETCH_L1(tp);
    PB_PREFETCH_L1(tp + 32);
  }

  // Regenerate the camera ray and its differentials (camera/mod.rs:212-271, ray.rs:107-112;
  // D15: the differential origins/directions stay in camera space).
  const float2 im = ld_stream(a.img + idx);
  float2 ln = make_float2(0.f, 0.f);
  if (a.lens) ln = ld_stream(a.lens + idx);
  f3 o, d, p_camera;
  camera_ray(cam, im.x, im.y, ln.x, ln.y, &o, &d, &p_camera);
  // The screen-space differentials (dpdx, dudx, ...) feed texture mappings only; a material whose
  // textures are all constant never reads them, so their value cannot influence the result and the
  // ~10 IEEE divisions / square roots behind them are skipped (mat_flags bit 0, set at upload).
  auto differentials = [&](DG& g) {
    const f3 dxc = mk3(cam.dx[0], cam.dx[1], cam.dx[2]), dyc = mk3(cam.dy[0], cam.dy[1], cam.dy[2]);
    f3 rxo = mk3(0.f, 0.f, 0.f), ryo = mk3(0.f, 0.f, 0.f);
    f3 rxd = normalize3(p_camera + dxc), ryd = normalize3(p_camera + dyc);
    const float s = cam.diff_scale;
    rxo = o + (rxo - o) * s;
    ryo = o + (ryo - o) * s;
    rxd = d + (rxd - d) * s;
    ryd = d + (ryd - d) * s;
    dg_compute_differentials(g, rxo, ryo, rxd, ryd);
  };

  // Intersection::get_bsdf
  uint32_t pr = sc.leaf_prim ? __ldg(&sc.leaf_prim[prim]) : prim;
  DG dg, dgs;
  uint32_t material;
  int32_t area_light = -1;
  [[maybe_unused]] bool shape_flip = false;  // reverse_orientation ^ transform_swaps_handedness (bump)
  if (pr & PB_LEAF_BIT) {
    const uint32_t si = pr & ~PB_LEAF_BIT;
    const pbrtb200_sphere80 sp = sc.spheres[si];
    material = sp.material;
    if constexpr (EXT) shape_flip = (sp.flip & 1u) != 0;
    dg = sphere_dg(sp, sc.sphere_o2w + 12ull * si, o, d, t_hit, hb1);
    if (__ldg(&sc.mat_flags[material]) & 1u) differentials(dg);
    dgs = dg;
  } else {
    const TriData td = load_tri(sc, pr);
    const pbrtb200_mesh m = sc.meshes[td.mesh];
    material = m.material;
    area_light = m.area_light;
    if constexpr (EXT) shape_flip = m.flip != 0;
    dg = tri_dg(td, o, d, t_hit, hb1, hb2, m.flip != 0);
    if (__ldg(&sc.mat_flags[material]) & 1u) differentials(dg);
    dgs = tri_shading_geometry(sc, td, m, dg);
  }
  ray_epsilon = t_hit * 5e-4f;  // mesh.rs:262, sphere.rs:180

  // Material::get_bsdf
  const pbrtb200_material mat = sc.materials[material];
  [[maybe_unused]] TexEnv env;
  if constexpr (EXT) {
    env.textures = sc.textures;
    env.mipmaps = sc.mipmaps;
    env.texels = sc.texels;
    // matte.rs:33-37 / plastic.rs:35-39: bump(tex, &dg_geom, &dg_shading)
    if (mat.bump > 0) dgs = bump_dg_(env, mat.bump - 1, dgs, dg.nn, shape_flip);
  }
  bs.nn = dgs.nn;                      // bsdf/mod.rs:70-86
  bs.tn = bs.sn = mk3(0.f, 0.f, 0.f);  // filled below for BxDFs that read the local frame
  bs.ng = dg.nn;
  bs.ks = mk3(0.f, 0.f, 0.f);
  bs.a = bs.b = 0.f;
  {
    f3 kd = shade_tex<IMG, EXT>(sc, env, mat.kd, dgs);
    bs.kd = mk3(rclampf(kd.x, 0.0f, PB_F32_MAX), rclampf(kd.y, 0.0f, PB_F32_MAX),
                rclampf(kd.z, 0.0f, PB_F32_MAX));
  }
  if (mat.kind == PBRTB200_MAT_MATTE) {  // matte.rs:30-51
    const float sig = rclampf(shade_tex<IMG, EXT>(sc, env, mat.sigma, dgs).x, 0.0f, 90.0f);
    if (sig == 0.0f) {
      bs.kind = 0;
    } else {  // orennayar.rs:16-29
      bs.kind = 1;
      const float sigma = sig * PB_PI / 180.0f;
      const float sigma2 = sigma * sigma;
      bs.a = 1.0f - (sigma2 / (2.0f * (sigma + 0.33f)));
      bs.b = 0.45f * sigma2 / (sigma2 + 0.09f);
    }
  } else {  // plastic.rs:32-55
    bs.kind = 2;
    const f3 ks = shade_tex<IMG, EXT>(sc, env, mat.ks, dgs);
    bs.ks = mk3(rclampf(ks.x, 0.0f, PB_F32_MAX), rclampf(ks.y, 0.0f, PB_F32_MAX),
                rclampf(ks.z, 0.0f, PB_F32_MAX));
    const float rough = shade_tex<IMG, EXT>(sc, env, mat.roughness, dgs).x;
    float e = 1.0f / rough;
    if (e > 1000.0f || isnan(e)) e = 1000.0f;  // microfacet.rs:18-24
    bs.a = e;
  }
  if (bs.kind != 0) {  // Lambertian::f ignores wo/wi (lambertian.rs:20-23): no frame needed
    bs.tn = normalize3(dgs.dpdu);
    bs.sn = cross3(bs.nn, bs.tn);
  }

  p = dgs.p;
  n = dgs.nn;
  wo = -d;

  // Emitted radiance at an emissive triangle (extension): L if n.w > 0.  Only the light's
  // index travels with the sample; the film / fold kernels add its radiance first (whitted.rs:46).
  if (area_light >= 0 && dot3(dg.nn, wo) > 0.0f) {
    le_bits = (uint32_t)area_light + 1u;
    const pbrtb200_light* al = sc.lights + area_light;
    if (isnan(al->intensity[0]) || isnan(al->intensity[1]) || isnan(al->intensity[2])) atomicAdd(a.nan_count, 1u);
  }
  }  // alive

  // Light-sample floats: 2 per area-light sample, drawn by raygen from the pixel's
  // stream after its camera-sample block, in (camera sample, light, light sample) order.
  const float2* lu = a.lightu ? a.lightu + idx * sc.area_sample_pairs : nullptr;

  uint32_t slot = 0;
  const uint32_t gslot0 = (uint32_t)(idx * a.slots);
  for (uint32_t li = 0; li < sc.n_lights; ++li) {
    const pbrtb200_light lt = sc.lights[li];
    const int ns = lt.kind == PBRTB200_LIGHT_AREA ? lt.num_samples : 1;
    for (int sidx = 0; sidx < ns; ++sidx, ++slot) {
      pbrtb200_ray32 vis;
      bool shadow = false, term_nan = false;
      if (alive) {
      f3 Li, wi;
      float pdf;
      if (lt.kind == PBRTB200_LIGHT_POINT) {  // point.rs:28-35 (D17: wi un-normalised)
        const f3 lp = mk3(lt.pos[0], lt.pos[1], lt.pos[2]);
        wi = lp - p;
        pdf = 1.0f;
        vis_segment(p, ray_epsilon, lp, 0.0f, &vis);
        Li = div3s(mk3(lt.intensity[0], lt.intensity[1], lt.intensity[2]), len2(wi));
      } else if (lt.kind == PBRTB200_LIGHT_SPOT) {  // spot.rs:37-65
        const f3 lp = mk3(lt.pos[0], lt.pos[1], lt.pos[2]);
        wi = normalize3(lp - p);
        pdf = 1.0f;
        vis_segment(p, ray_epsilon, lp, 0.0f, &vis);
        const f3 wl = xf_vec(lt.w2l, -wi);
        const float cos_theta = wl.z;
        float fall;
        if (cos_theta < lt.cos_total_width)
          fall = 0.0f;
        else if (cos_theta > lt.cos_falloff_start)
          fall = 1.0f;
        else {
          const float delta =
              (cos_theta - lt.cos_total_width) / (lt.cos_falloff_start - lt.cos_total_width);
          fall = delta * delta * delta * delta;
        }
        const f3 I = mk3(lt.intensity[0], lt.intensity[1], lt.intensity[2]) * fall;
        Li = div3s(I, len2(wi));
      } else {
        // Diffuse area light over emissive triangles (extension; pbrt-v2 semantics, A13)
        const float2 uu = ld_stream(lu++);
        const float u1 = uu.x, u2 = uu.y;
        uint32_t k = 0;
        while (k + 1 < lt.n_tris && u1 >= a.area_tris[lt.first_tri + k].cdf_hi) ++k;
        const DAreaTri at = a.area_tris[lt.first_tri + k];
        const float u1p = (u1 - at.cdf_lo) / (at.cdf_hi - at.cdf_lo);
        const float su = sqrtf(u1p);
        const float b0 = 1.0f - su, b1 = u2 * su;
        const f3 ps = b0 * mk3(at.p1[0], at.p1[1], at.p1[2]) + b1 * mk3(at.p2[0], at.p2[1], at.p2[2]) +
                      (1.0f - b0 - b1) * mk3(at.p3[0], at.p3[1], at.p3[2]);
        wi = normalize3(ps - p);
        const float cos_l = dot3(mk3(at.nn[0], at.nn[1], at.nn[2]), -wi);
        const float d2 = len2(ps - p);
        vis_segment(p, ray_epsilon, ps, 1e-3f, &vis);
        if (!(cos_l > 0.0f)) {
          Li = mk3(0.f, 0.f, 0.f);
          pdf = 0.0f;
        } else {
          Li = mk3(lt.intensity[0], lt.intensity[1], lt.intensity[2]);
          pdf = d2 / (fabsf(cos_l) * lt.total_area);
        }
      }
      f3 c = mk3(0.f, 0.f, 0.f);
      if (!(is_black(Li) || pdf == 0.0f)) {  // whitted.rs:55
        const f3 f = bsdf_f(bs, wo, wi, a.strict_flags != 0);
        if (!is_black(f)) {
          // whitted.rs:60-63 with T = 1 (D3): ((f * li) * |wi.n|) * T / pdf
          c = div3s(mul3(mul3(f, Li) * fabsf(dot3(wi, n)), mk3(1.0f, 1.0f, 1.0f)), pdf);
          shadow = true;
        }
      }
      term_nan = isnan(c.x) || isnan(c.y) || isnan(c.z);
      st_stream(terms + slot, make_float4(c.x, c.y, c.z, slot == 0u ? __uint_as_float(le_bits) : 0.f));
      }  // alive
      // block-aggregated push: ballot per warp, one global atomic per block and slot
      const unsigned sm = __ballot_sync(0xffffffffu, shadow);
      const uint32_t* wc = s_warp_cnt[slot & 1u];
      if (lane == 0) s_warp_cnt[slot & 1u][warp] = __popc(sm);
      __syncthreads();
      if (threadIdx.x == 0) {
        const uint32_t t = wc[0] + wc[1] + wc[2] + wc[3];
        s_block_base[slot & 1u] = t ? atomicAdd(a.sq_count, t) : 0u;
      }
      __syncthreads();
      if (shadow) {
        uint32_t q = s_block_base[slot & 1u] + (uint32_t)__popc(sm & ((1u << lane) - 1u));
        for (int w = 0; w < warp; ++w) q += wc[w];
        float4* rq = reinterpret_cast<float4*>(a.sq_rays + q);
        st_stream(rq, make_float4(vis.o[0], vis.o[1], vis.o[2], vis.mint));
        st_stream(rq + 1, make_float4(vis.d[0], vis.d[1], vis.d[2], vis.maxt));
        st_stream(a.sq_slots + q, (gslot0 + slot) | (term_nan ? PB_SQ_NAN : 0u) | ((slot == 0u && le_bits) ? PB_SQ_KEEPW : 0u));
      }
    }
  }
}

// Upload-time helper: fill DAreaTri records from ordered-primitive indices (extension, A13).
__global__ void k_area_tri_setup(const DScene sc, const uint32_t* __restrict__ prims, uint32_t n,
                                 DAreaTri* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t prim = prims[i];
  const uint32_t pr = sc.leaf_prim ? sc.leaf_prim[prim] : prim;
  const TriData td = load_tri(sc, pr & ~PB_LEAF_BIT);
  const pbrtb200_mesh m = sc.meshes[td.mesh];
  const DG dg = tri_dg(td, mk3(0, 0, 0), mk3(0, 0, 1), 0.f, 0.f, 0.f, m.flip != 0);
  DAreaTri t;
  t.p1[0] = td.p1.x; t.p1[1] = td.p1.y; t.p1[2] = td.p1.z;
  t.p2[0] = td.p2.x; t.p2[1] = td.p2.y; t.p2[2] = td.p2.z;
  t.p3[0] = td.p3.x; t.p3[1] = td.p3.y; t.p3[2] = td.p3.z;
  t.nn[0] = dg.nn.x; t.nn[1] = dg.nn.y; t.nn[2] = dg.nn.z;
  t.area = 0.5f * len3(cross3(td.p2 - td.p1, td.p3 - td.p1));  // mesh.rs:100-103
  t.cdf_lo = t.cdf_hi = t.pad = 0.f;
  out[i] = t;
}
