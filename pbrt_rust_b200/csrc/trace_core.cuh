// One ray through the pair-node BVH (sm_100a): the quadric intersection tests, the traversal loop
// and its dispatch on the ray's direction.  See trace.cuh for the equivalence with the reference's
// test-at-pop order.  The functions only touch memory through __ldg and the stack pointers they are
// handed, so the same source compiles as host code (PB_HOST_CHECK, tests/devsrc/): the CPU test-suite
// runs this very traversal over the host mirror's flattened scenes against the oracle.
#pragma once
#include "scene.cuh"
#include "trace_math.cuh"

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

