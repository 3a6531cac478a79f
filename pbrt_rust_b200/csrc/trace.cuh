// Closest-hit / any-hit BVH traversal kernels (sm_100a).
//
// Replaces, per ray: BVHAccelerator::intersect (src/primitive/aggregates/bvh.rs:375-422),
// BBox::intersect (src/bbox.rs:185-209), Triangle::get_intersection_point (src/shape/mesh.rs:41-72),
// Sphere::get_intersection_point (src/shape/sphere.rs:46-107) and GeometricPrimitive::intersect's
// ray.maxt update (src/primitive/geometric.rs:62-73).
//
// Equivalence with the reference's test-at-pop order (proved in DESIGN.md §"Traversal"): the
// reference pops a node and tests its box against the LIVE [mint,maxt]; the test passes iff
// T0 <= F and T0 <= maxt_now with T0 = max(mint, near_xyz), F = min(far_xyz) (both independent of
// maxt).  We evaluate both children's boxes when visiting the parent, push the far child together
// with its T0 only if it passes then, and re-check T0 <= maxt when it is popped.  Because maxt
// only shrinks, the set and ORDER of visited leaves — hence every accepted hit, including the
// "equal t replaces" tie rule (mesh.rs:71, bvh.rs:401-404) — is identical.
#pragma once
#include "scene.cuh"
#include "trace_math.cuh"  // slab_test, slab_test_finite, tri_hit, camera_ray

PB_DEV float4 ldg4(const float4* p) { return __ldg(p); }

// sphere.rs:46-107 (+ the world->object ray transform of sphere.rs:137)
PB_DEV bool sphere_hit(const pbrtb200_sphere80* __restrict__ sp, f3 ow, f3 dw, float mint,
                       float maxt, float* t_out, float* phi_out) {
  float m[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) m[i] = __ldg(&sp->w2o[i]);
  const float radius = __ldg(&sp->radius), z_min = __ldg(&sp->z_min), z_max = __ldg(&sp->z_max),
              phi_max = __ldg(&sp->phi_max);
  const uint32_t kind = (__ldg(&sp->flip) >> PBRTB200_QUADRIC_KIND_SHIFT) & 3u;
  f3 o = xf_pt(m, ow), d = xf_vec(m, dw);
  if (kind == PBRTB200_QUADRIC_DISK) {  // disk.rs:37-71 (z_min = height, theta_min = inner radius)
    if (fabsf(d.z) < 1e-6f) return false;
    const float t_hit = (z_min - o.z) / d.z;
    if (t_hit < mint || t_hit > maxt) return false;
    const f3 p_hit = o + (d * t_hit);
    const float dist2 = p_hit.x * p_hit.x + p_hit.y * p_hit.y;
    const float inner = __ldg(&sp->theta_min);
    if (dist2 > (radius * radius) || dist2 < (inner * inner)) return false;
    const float a = atan2f(p_hit.y, p_hit.x);
    const float phi = a < 0.0f ? a + 2.0f * PB_PI : a;
    if (phi > phi_max) return false;
    *t_out = t_hit;
    *phi_out = phi;
    return true;
  }
  const bool cyl = kind == PBRTB200_QUADRIC_CYLINDER;  // cylinder.rs:40-100 shares sphere.rs's flow
  float a = cyl ? d.x * d.x + d.y * d.y : len2(d);
  float b = cyl ? 2.0f * (d.x * o.x + d.y * o.y) : 2.0f * dot3(d, o);
  float c = (cyl ? o.x * o.x + o.y * o.y : len2(o)) - radius * radius;
  float t0, t1;
  if (!quadratic_(a, b, c, &t0, &t1)) return false;
  if (t0 > maxt || t1 < mint) return false;
  float t_hit = t0;
  if (t0 < mint) {
    t_hit = t1;
    if (t_hit > maxt) return false;
  }
  f3 h = o + (d * t_hit);
  if (h.x == 0.0f && h.y == 0.0f) h.x = 1e-5f * radius;
  float ang = atan2f(h.y, h.x);
  if (ang < 0.0f) ang += 2.0f * PB_PI;
  auto clipped = [&](f3 hp, float an) {  // sphere.rs:84-88 / cylinder.rs:78-80
    return cyl ? (hp.z < z_min || hp.z > z_max || an > phi_max)
               : ((hp.z > -radius && hp.z < z_min) || (hp.z < radius && hp.z > z_max) || (an > phi_max));
  };
  if (clipped(h, ang)) {
    if (t_hit == t1) return false;
    if (t1 > maxt) return false;
    t_hit = t1;
    h = o + (d * t_hit);
    if (h.x == 0.0f && h.y == 0.0f) h.x = 1e-5f * radius;
    ang = atan2f(h.y, h.x);
    if (ang < 0.0f) ang += 2.0f * PB_PI;
    if (clipped(h, ang)) return false;
  }
  *t_out = t_hit;
  *phi_out = ang;
  return true;
}

struct TraceResult {
  uint32_t prim;
  float t, b1, b2;
  bool overflow;
};

#define PB_DONE 0xFFFFFFFFu  // traversal finished (has the leaf bit set, so it leaves the node loop)

// One ray through the pair-node BVH.  s_ref / s_t0 point at this thread's column of the shared
// stack (stride PB_TRACE_THREADS).  ANY: stop at the first accepted hit (VisibilityTester).
// FINITE: every component of 1/d is finite (the common case; selects the cheaper slab test).
// MODE selects the SIMT loop shape (both visit the same leaves in the same order):
//   0  if-if        : each iteration a lane does one node step OR one leaf        (best for any-hit)
//   1  while-while  : lanes run node steps until every lane of the warp holds a leaf (best closest)
//   2, 3 (ANY only) : shapes 0 / 1 without the near/far child ordering.  An any-hit query is a
//                     boolean over the set of leaves whose boxes pass; no accepted hit shrinks maxt
//                     before it returns, so that set does not depend on the visiting order and the
//                     axis decode + selects of bvh.rs:409-415 buy nothing for unoccluded rays.
// Measured and dropped (profiles/r01_notes.md): speculative postponed-leaf traversal (no gain) and a
// persistent kernel with per-lane ray refill (-30..-70 %: refilled lanes lose ray coherence), and a
// warp-cooperative any-hit kernel with subtree stealing between lanes (-18 %).
template <bool ANY, bool SPH, bool MULTI, bool FINITE, int MODE>
PB_DEV TraceResult traverse(const DScene& sc, f3 o, f3 d, f3 inv, float mint, float maxt,
                            uint32_t* s_ref, float* s_t0) {
  constexpr int stride = PB_TRACE_THREADS;
  constexpr bool UNORDERED = ANY && MODE >= 2;
  constexpr int SHAPE = MODE & 1;
  TraceResult res;
  res.prim = PBRTB200_MISS;
  res.t = 0.f;
  res.b1 = 0.f;
  res.b2 = 0.f;
  res.overflow = false;
  // bvh.rs:382-383
  const bool neg0 = inv.x < 0.0f, neg1 = inv.y < 0.0f, neg2 = inv.z < 0.0f;
  uint32_t l_ref[PB_LM_STACK];
  float l_t0[PB_LM_STACK];
  int sp = 0;
  auto box = [&](float ax, float ay, float az, float bx, float by, float bz, float* T0) {
    return FINITE ? slab_test_finite(ax, ay, az, bx, by, bz, o, inv, mint, maxt, T0)
                  : slab_test(ax, ay, az, bx, by, bz, o, inv, mint, maxt, T0);
  };
  // pop the next stack entry that still passes the reference's box test at pop (live maxt).
  // ANY: an any-hit query returns at its first accepted hit, so maxt never shrinks while entries
  // are on the stack; every pushed entry passed with this very maxt and T0 need not be kept.
  auto pop = [&]() -> uint32_t {
    while (sp > 0) {
      --sp;
      uint32_t r;
      float t0 = 0.f;
      if (sp < PB_SM_STACK) {
        r = s_ref[sp * stride];
        if (!ANY) t0 = s_t0[sp * stride];
      } else {
        r = l_ref[sp - PB_SM_STACK];
        if (!ANY) t0 = l_t0[sp - PB_SM_STACK];
      }
      if (ANY || !(t0 > maxt)) return r;
    }
    return PB_DONE;
  };
  // one inner-node step: both children's boxes, descend near / push far / pop
  auto node_step = [&](uint32_t cur) -> uint32_t {
    const float4* n = sc.nodes + 4ull * cur;
    const float4 q0 = ldg4(n), q1 = ldg4(n + 1), q2 = ldg4(n + 2), q3 = ldg4(n + 3);
    float T00, T01;
    const bool h0 = box(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &T00);
    const bool h1 = box(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &T01);
    const uint32_t r0 = __float_as_uint(q3.x), r1 = __float_as_uint(q3.y);
    if (h0 & h1) {
      // bvh.rs:409-415: dir_is_neg[axis] -> the second child is visited first
      bool neg = false;
      if (!UNORDERED) {
        const uint32_t axis = __float_as_uint(q3.w);
        neg = axis == 0 ? neg0 : (axis == 1 ? neg1 : neg2);
      }
      const uint32_t far_ref = neg ? r0 : r1;
      const float far_t0 = neg ? T00 : T01;
      if (sp < PB_SM_STACK) {
        s_ref[sp * stride] = far_ref;
        if (!ANY) s_t0[sp * stride] = far_t0;
      } else if (sp < PBRTB200_STACK_DEPTH) {
        l_ref[sp - PB_SM_STACK] = far_ref;
        if (!ANY) l_t0[sp - PB_SM_STACK] = far_t0;
      } else {
        res.overflow = true;
        return PB_DONE;
      }
      ++sp;
      return neg ? r1 : r0;
    }
    if (h0 | h1) return h0 ? r0 : r1;
    return pop();
  };
  // bvh.rs:398-405: every primitive of the leaf in order; the last accepted hit wins.
  // Returns true when an ANY-hit query is answered.
  auto leaf = [&](uint32_t ref) -> bool {
    const uint32_t off = ref & PB_LEAF_OFF_MASK;
    uint32_t cnt = 1u;
    if (MULTI) {
      cnt = ((ref >> PB_LEAF_CNT_SHIFT) & 0xFu) + 1u;  // 1..15 inline; 16 = look it up
      if (cnt == 16u) cnt = (uint32_t)__ldg(&sc.leaf_count[off]);
    }
    for (uint32_t i = 0; i < cnt; ++i) {
      const uint32_t pi = off + i;
      uint32_t pr = SPH ? __ldg(&sc.leaf_prim[pi]) : pi;
      bool hit;
      float t, b1, b2 = 0.f;
      if (SPH && (pr & PB_LEAF_BIT)) {
        hit = sphere_hit(sc.spheres + (pr & ~PB_LEAF_BIT), o, d, mint, maxt, &t, &b1);
      } else {
        const float4* tp = sc.tris + 3ull * pr;
        const float4 a = ldg4(tp), b = ldg4(tp + 1), c = ldg4(tp + 2);
        hit = tri_hit(mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), mk3(c.x, c.y, c.z), o, d, mint,
                      maxt, &t, &b1, &b2);
      }
      if (hit) {
        maxt = t;  // geometric.rs:64
        res.prim = pi;
        res.t = t;
        res.b1 = b1;
        res.b2 = b2;
        if (ANY) return true;
      }
    }
    return false;
  };

  float T0root;
  if (!box(sc.root_bmin[0], sc.root_bmin[1], sc.root_bmin[2], sc.root_bmax[0], sc.root_bmax[1],
           sc.root_bmax[2], &T0root))
    return res;
  uint32_t cur = sc.root_ref;
  if (SHAPE == 0) {
    while (cur != PB_DONE) {
      if (!(cur & PB_LEAF_BIT)) {
        cur = node_step(cur);
      } else {
        if (leaf(cur)) return res;
        cur = pop();
      }
    }
  } else {
    for (;;) {
      while (!(cur & PB_LEAF_BIT)) cur = node_step(cur);
      if (cur == PB_DONE) break;
      if (leaf(cur)) return res;
      cur = pop();
    }
  }
  return res;
}

template <bool ANY, bool SPH, bool MULTI, int MODE>
PB_DEV TraceResult trace_ray(const DScene& sc, f3 o, f3 d, float mint, float maxt, uint32_t* s_ref,
                             float* s_t0) {
  const f3 inv = mk3(1.f / d.x, 1.f / d.y, 1.f / d.z);  // bvh.rs:382
  const float big = fmaxf(fmaxf(fabsf(inv.x), fabsf(inv.y)), fabsf(inv.z));
  // (NaN-propagating test: a NaN or infinite component takes the exact-compare path)
  if (big < __int_as_float(0x7f800000) && inv.x == inv.x && inv.y == inv.y && inv.z == inv.z)
    return traverse<ANY, SPH, MULTI, true, MODE>(sc, o, d, inv, mint, maxt, s_ref, s_t0);
  return traverse<ANY, SPH, MULTI, false, 0>(sc, o, d, inv, mint, maxt, s_ref, s_t0);
}

// ---- kernels ---------------------------------------------------------------------------------
// Persistent grid: each warp pulls 32-ray packets from a global counter (warp-aggregated: one
// atomic per warp), so slow packets do not stall a whole block's worth of queued work.

struct TraceArgs {
  const pbrtb200_ray32* __restrict__ rays;  // SRC 0
  const float2* __restrict__ img;           // SRC 1: camera samples (image_x, image_y)
  const float2* __restrict__ lens;          // SRC 1, may be NULL (lens_radius == 0)
  pbrtb200_hit16* __restrict__ hits;        // closest-hit output (may be NULL for ANY)
  uint8_t* __restrict__ occluded;           // ANY output (API hook), may be NULL
  float4* __restrict__ contrib;             // ANY in the render pipeline: zero slot if occluded
  const uint32_t* __restrict__ slots;       // ANY in the render pipeline: slot index per ray
  const uint32_t* __restrict__ n_dyn;       // if non-NULL the ray count is read from device memory
  unsigned long long* shadow_total;         // if non-NULL: += ray count (stats)
  uint64_t n;
  unsigned long long* counter;  // work counter, zeroed by the host before launch
  uint32_t* flags;              // bit0: stack overflow happened
  uint32_t batch;               // 32-ray packets a warp claims per atomic (>= 1)
  // SRC 1, tiled renders: a halo pixel whose samples all stay inside their own pixel (edge flag 0,
  // set by raygen with add_sample's own arithmetic) cannot reach a pixel this call owns, so its
  // rays are not traced (hit = MISS).  With the box filter that is practically every halo pixel.
  const DPixel* __restrict__ pixels;   // may be NULL: trace everything
  const uint32_t* __restrict__ edge;
  uint32_t spp;
};

template <bool ANY, bool SPH, bool MULTI, int SRC, int MODE>
__global__ void __launch_bounds__(PB_TRACE_THREADS)
k_trace(const DScene sc, const DCamera cam, const TraceArgs a) {
  __shared__ uint32_t sh_ref[PB_SM_STACK * PB_TRACE_THREADS];
  __shared__ float sh_t0[ANY ? 1 : PB_SM_STACK * PB_TRACE_THREADS];  // any-hit keeps no T0
  uint32_t* s_ref = sh_ref + threadIdx.x;
  float* s_t0 = ANY ? sh_t0 : sh_t0 + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const uint64_t n = a.n_dyn ? (uint64_t)(*a.n_dyn) : a.n;
  if (a.shadow_total && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(a.shadow_total, (unsigned long long)n);
  const unsigned long long claim = 32ull * a.batch;
  for (;;) {
    unsigned long long base0 = 0;
    if (lane == 0) base0 = atomicAdd(a.counter, claim);
    base0 = __shfl_sync(0xffffffffu, base0, 0);
    if (base0 >= n) break;  // warp-uniform exit
   for (unsigned long long base = base0; base < base0 + claim && base < n; base += 32ull) {
    const uint64_t idx = base + lane;
    if (idx < n) {
      f3 o, d;
      float mint, maxt;
      if (SRC == 0) {
        const float4* rp = reinterpret_cast<const float4*>(a.rays + idx);
        const float4 r0 = ldg4(rp), r1 = ldg4(rp + 1);
        o = mk3(r0.x, r0.y, r0.z);
        mint = r0.w;
        d = mk3(r1.x, r1.y, r1.z);
        maxt = r1.w;
      } else {
        if (a.pixels) {
          const uint64_t pix = idx / a.spp;
          if ((__ldg(&a.pixels[pix].task) & PB_PIXEL_HALO_BIT) && __ldg(&a.edge[pix]) == 0u) {
            reinterpret_cast<float4*>(a.hits)[idx] = make_float4(__uint_as_float(PBRTB200_MISS), 0.f, 0.f, 0.f);
            continue;
          }
        }
        const float2 im = __ldg(a.img + idx);
        float2 ln = make_float2(0.f, 0.f);
        if (a.lens) ln = __ldg(a.lens + idx);
        camera_ray(cam, im.x, im.y, ln.x, ln.y, &o, &d, nullptr);
        mint = 0.0f;         // ray.rs:30-38 Ray::new_with(.., start = 0)
        maxt = PB_F32_MAX;
      }
      TraceResult r = trace_ray<ANY, SPH, MULTI, MODE>(sc, o, d, mint, maxt, s_ref, s_t0);
      if (r.overflow) atomicOr(a.flags, 1u);
      if (ANY) {
        const bool occ = r.prim != PBRTB200_MISS;
        if (a.occluded) a.occluded[idx] = occ ? 1 : 0;
        if (a.contrib && occ) a.contrib[__ldg(a.slots + idx)] = make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        const uint64_t oi = idx;
        float4 h;
        h.x = __uint_as_float(r.prim);
        h.y = r.t;
        h.z = r.b1;
        h.w = r.b2;
        reinterpret_cast<float4*>(a.hits)[oi] = h;
      }
    }
   }
  }
}


