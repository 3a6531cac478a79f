// One ray through the pair-node BVH (sm_100a): the quadric intersection tests, the traversal loop
// and its dispatch on the ray's direction.  See trace.cuh for the equivalence with the reference's
// test-at-pop order.  The functions only touch memory through __ldg and the stack pointers they are
// handed, so the same source compiles as host code (PB_HOST_CHECK, tests/devsrc/): the CPU test-suite
// runs this very traversal over the host mirror's flattened scenes against the oracle.
#pragma once
#include "scene.cuh"
#include "trace_math.cuh"

PB_DEV float4 ldg4(const float4* p) { return __ldg(p); }
// Two consecutive float4 (32-byte aligned), optionally in ONE 256-bit load (sm_100: LDG.E.256,
// -DPB_LDG256=1).  Measured (profiles/r02_notes.md §10): L1 requests and fetched sectors of the
// closest-hit kernel halve (145 M -> 80 M, 220 M -> 126 M), instructions -2.3 %, but the L1 data pipe
// stays at 79 % — it is busy delivering the 64 bytes per lane, not taking requests — and the kernel
// times do not move (2.926 -> 2.919 ms, any-hit 3.051 -> 3.075 ms).  Off by default.
#ifndef PB_LDG256
#define PB_LDG256 0
#endif
PB_DEV void ldg8(const float4* p, float4* a, float4* b) {
#if PB_LDG256 && !defined(PB_HOST_CHECK)
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(a->x), "=f"(a->y), "=f"(a->z), "=f"(a->w), "=f"(b->x), "=f"(b->y), "=f"(b->z), "=f"(b->w)
      : "l"(p));
#else
  *a = ldg4(p);
  *b = ldg4(p + 1);
#endif
}

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
  uint32_t prim;  // PBRTB200_MISS, PB_OVERFLOW (the stack overflowed: result invalid) or the hit
  float t, b1, b2;
  uint32_t steps;  // COUNT instantiations only: node steps + primitive tests (cost probe, api.cu)
};
#define PB_OVERFLOW 0xFFFFFFFEu

#define PB_DONE 0xFFFFFFFFu  // traversal finished (has the leaf bit set, so it leaves the node loop)
#define PB_DONE_OVF 0xFFFFFFFEu  // ... because the stack overflowed (no leaf ref: n_prims < 2^27 - 16)

// Per-ray constants of the box tests.
struct RayBox {
  f3 o, inv;    // exact tests: (plane - o) * inv, bbox.rs:190-193
  f3 ncn, ncf;  // BOX 3 only: t_near = fma(plane, inv, ncn), t_far = fma(plane, inv, ncf)
};

// The box test of one child, selected at compile time.
//   BOX 0  bbox.rs:185-209 literally (compare + swap): reproduces the reference's NaN behaviour for
//          rays with a zero direction component.
//   BOX 1  the same values through min / max (every 1/d finite, box coordinates finite).
//   BOX 2  BOX 1 specialised for the sign of 1/d per axis (AX, below): with bmin <= bmax the
//          products (bmin - o) * inv and (bmax - o) * inv are ordered by the sign of inv alone
//          (rounding is monotonic), so the swap of bbox.rs:194-196 is decided at compile time and the
//          six per-axis min / max disappear.  Same values as BOX 1, bit for bit.
//   BOX 3  CONSERVATIVE test for inner nodes: one FFMA per plane, t = plane * inv - (o * inv -/+ e),
//          where e bounds the rounding difference to the reference's (plane - o) * inv for every
//          plane inside the scene bounds (trace_ray computes it per ray).  It passes whenever the
//          reference's test passes (T0c <= T0, Fc >= F), possibly more often.  By the containment
//          lemma (DESIGN.md "Traversal") which leaves are visited, and in which order, is decided by
//          each leaf's own exact test, which BOX 3 applies in the leaf phase to every primitive that
//          reports a hit; inner tests may be any superset.
// AX (BOX 2 / 3): per axis a = x, y, z two bits, (AX >> 2a) & 3 = 0: every ray of the warp has
// 1/d[a] >= 0, 1: every ray has 1/d[a] < 0, 2: the warp mixes signs on this axis (that axis keeps
// the min / max form; shadow rays towards a light overhead mix x and z in most warps).
template <int BOX, int AX>
PB_DEV void axis_t(float lo, float hi, float o, float inv, float ncn, float ncf, int a, float* tn, float* tf) {
  const int s = (AX >> (2 * a)) & 3;
  if (BOX == 2) {
    if (s == 2) {
      const float ta = (lo - o) * inv, tb = (hi - o) * inv;
      *tn = fminf(ta, tb);
      *tf = fmaxf(ta, tb);
    } else {
      *tn = ((s ? hi : lo) - o) * inv;
      *tf = ((s ? lo : hi) - o) * inv;
    }
  } else {
    if (s == 2) {
      *tn = fminf(fmaf(lo, inv, ncn), fmaf(hi, inv, ncn));
      *tf = fmaxf(fmaf(lo, inv, ncf), fmaf(hi, inv, ncf));
    } else {
      *tn = fmaf(s ? hi : lo, inv, ncn);
      *tf = fmaf(s ? lo : hi, inv, ncf);
    }
  }
}
template <int BOX, int AX>
PB_DEV bool child_box(const RayBox& rb, float ax, float ay, float az, float bx, float by, float bz,
                      float mint, float maxt, float* T0) {
  if (BOX == 0) return slab_test(ax, ay, az, bx, by, bz, rb.o, rb.inv, mint, maxt, T0);
  if (BOX == 1) return slab_test_finite(ax, ay, az, bx, by, bz, rb.o, rb.inv, mint, maxt, T0);
  float tnx, tny, tnz, tfx, tfy, tfz;
  axis_t<BOX, AX>(ax, bx, rb.o.x, rb.inv.x, rb.ncn.x, rb.ncf.x, 0, &tnx, &tfx);
  axis_t<BOX, AX>(ay, by, rb.o.y, rb.inv.y, rb.ncn.y, rb.ncf.y, 1, &tny, &tfy);
  axis_t<BOX, AX>(az, bz, rb.o.z, rb.inv.z, rb.ncn.z, rb.ncf.z, 2, &tnz, &tfz);
  const float t0 = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, mint));
  const float t1 = fminf(fminf(tfx, tfy), fminf(tfz, maxt));
  *T0 = t0;
  return !(t0 > t1);
}

// The traversal stack of one thread: PB_SM_STACK entries in shared memory (conflict-free
// [depth][lane] layout: entry k of a thread lives PB_TRACE_THREADS words after entry k - 1), deeper
// entries in local memory.  An entry is a child ref, plus (closest hit only) the child's entry
// distance T0 for the re-check at pop.  On the device the stack pointer IS the shared-memory byte
// address of the next free entry, so a push / pop is one add and one or two STS / LDS with an
// immediate offset — no index arithmetic (the kernels are bound by instruction issue).
template <bool ANY>
struct TStack {
  static constexpr int SMD = PB_SM_STACK_OF(ANY);                 // entries in shared memory
  static constexpr int LMD = PBRTB200_STACK_DEPTH - SMD;          // entries in local memory
#ifdef PB_HOST_CHECK
  uint32_t* s_ref;
  float* s_t0;
  int sp;
  PB_DEV void init(uint32_t* ref, float* t0) {
    s_ref = ref;
    s_t0 = t0;
    sp = 0;
  }
  PB_DEV bool empty() const { return sp == 0; }
  PB_DEV int depth() const { return sp; }
  PB_DEV bool in_shared() const { return sp < SMD; }
  PB_DEV void put(uint32_t r, float t0) {
    s_ref[sp * PB_TRACE_THREADS] = r;
    if (!ANY) s_t0[sp * PB_TRACE_THREADS] = t0;
  }
  PB_DEV void get(uint32_t* r, float* t0) const {
    *r = s_ref[sp * PB_TRACE_THREADS];
    if (!ANY) *t0 = s_t0[sp * PB_TRACE_THREADS];
  }
  PB_DEV void up() { ++sp; }
  PB_DEV void down() { --sp; }
#else
  static constexpr uint32_t kStep = 4u * PB_TRACE_THREADS;           // bytes between entries
  static constexpr uint32_t kT0 = 4u * (SMD > 0 ? SMD : 1) * PB_TRACE_THREADS;  // T0 array follows the refs
  uint32_t top;  // shared-window byte address of the next free entry (while sp <= PB_SM_STACK)
  int sp;        // entries on the stack (limit checks only; the address is never derived from it)
  PB_DEV void init(uint32_t* ref, float*) {  // (s_t0 == s_ref + SMD * PB_TRACE_THREADS)
    top = SMD > 0 ? (uint32_t)__cvta_generic_to_shared(ref) : 0u;
    sp = 0;
  }
  PB_DEV bool empty() const { return sp == 0; }
  PB_DEV int depth() const { return sp; }
  PB_DEV bool in_shared() const { return sp < SMD; }
  PB_DEV void put(uint32_t r, float t0) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(top), "r"(r) : "memory");
    if (!ANY) asm volatile("st.shared.f32 [%0+%2], %1;" ::"r"(top), "f"(t0), "n"(kT0) : "memory");
  }
  PB_DEV void get(uint32_t* r, float* t0) const {
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(*r) : "r"(top) : "memory");
    if (!ANY) asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(*t0) : "r"(top), "n"(kT0) : "memory");
  }
  PB_DEV void up() {
    if (SMD > 0) top += kStep;
    ++sp;
  }
  PB_DEV void down() {
    if (SMD > 0) top -= kStep;
    --sp;
  }
#endif
};

// One ray through the pair-node BVH.  s_ref / s_t0 point at this thread's column of the shared
// stack (stride PB_TRACE_THREADS; s_t0 == s_ref + PB_SM_STACK_OF(ANY) * PB_TRACE_THREADS).  ANY: stop at
// the first accepted hit (VisibilityTester).
// BOX / AX: the box test (child_box).  MODE selects the SIMT loop shape (all visit the same leaves
// in the same order):
//   0  if-if        : each iteration a lane does one node step OR one leaf        (best for any-hit)
//   1  while-while  : lanes run node steps until every lane of the warp holds a leaf (best closest)
//   4 (ANY only)    : unordered, with postponed leaves (see below)
//   2, 3 (ANY only) : shapes 0 / 1 without the near/far child ordering.  An any-hit query is a
//                     boolean over the set of leaves whose boxes pass; no accepted hit shrinks maxt
//                     before it returns, so that set does not depend on the visiting order and the
//                     axis decode + selects of bvh.rs:409-415 buy nothing for unoccluded rays.
// Measured and dropped (profiles/r01_notes.md): speculative postponed-leaf traversal (no gain) and a
// persistent kernel with per-lane ray refill (-30..-70 %: refilled lanes lose ray coherence), and a
// warp-cooperative any-hit kernel with subtree stealing between lanes (-18 %).
template <bool ANY, bool SPH, bool MULTI, int BOX, int MODE, int AX, bool COUNT = false>
PB_DEV TraceResult traverse(const DScene& sc, f3 o, f3 d, const RayBox& rb, float mint, float maxt,
                            uint32_t* s_ref, float* s_t0) {
  constexpr bool UNORDERED = ANY && MODE >= 2;
  constexpr int SHAPE = MODE & 1;
  constexpr bool LEAF_EXACT = BOX == 3;  // inner tests are conservative: exact leaf test in the leaf phase
  constexpr int SMD = TStack<ANY>::SMD, LMD = TStack<ANY>::LMD;
  constexpr bool SM = SMD > 0;
  TraceResult res;
  res.prim = PBRTB200_MISS;
  res.t = 0.f;
  res.b1 = 0.f;
  res.b2 = 0.f;
  res.steps = 0u;
  // bvh.rs:382-383, as a mask over the q3.z axis bit of a node
  const uint32_t oct = (rb.inv.x < 0.0f ? 1u : 0u) | (rb.inv.y < 0.0f ? 2u : 0u) | (rb.inv.z < 0.0f ? 4u : 0u);
  uint32_t l_ref[LMD];
  float l_t0[ANY ? 1 : LMD];
  TStack<ANY> st;
  st.init(s_ref, s_t0);
  auto box = [&](float ax, float ay, float az, float bx, float by, float bz, float* T0) {
    return child_box<BOX, AX>(rb, ax, ay, az, bx, by, bz, mint, maxt, T0);
  };
  // pop the next stack entry that still passes the reference's box test at pop (live maxt).
  // ANY: an any-hit query returns at its first accepted hit, so maxt never shrinks while entries
  // are on the stack; every pushed entry passed with this very maxt and T0 need not be kept.
  // (BOX 3: the kept T0 is the conservative one, T0c <= T0: the re-check stays a superset.)
  auto pop = [&]() -> uint32_t {
    while (!st.empty()) {
      st.down();
      uint32_t r;
      float t0 = 0.f;
      if (SM && st.in_shared()) {
        st.get(&r, &t0);
      } else {
        const int k = st.depth() - SMD;
        r = l_ref[k];
        if (!ANY) t0 = l_t0[k];
      }
      if (ANY || !(t0 > maxt)) return r;
    }
    return PB_DONE;
  };
  // one inner-node step: both children's boxes, descend near / push far / pop
  auto node_step = [&](uint32_t cur) -> uint32_t {
    if (COUNT) ++res.steps;
    const float4* n = sc.nodes + 4ull * cur;
    float4 q0, q1, q2, q3;
    ldg8(n, &q0, &q1);
    ldg8(n + 2, &q2, &q3);
    float T00, T01;
    const bool h0 = box(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &T00);
    const bool h1 = box(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &T01);
    const uint32_t r0 = __float_as_uint(q3.x), r1 = __float_as_uint(q3.y);
    if (h0) {
      if (!h1) return r0;
      // both pass.  bvh.rs:409-415: dir_is_neg[axis] -> the second child is visited first
      bool neg = false;
      if (!UNORDERED) neg = (__float_as_uint(q3.z) & oct) != 0u;  // q3.z = 1 << axis
      const uint32_t far_ref = neg ? r0 : r1;
      const float far_t0 = neg ? T00 : T01;
      if (SM && st.in_shared()) {
        st.put(far_ref, far_t0);
      } else {
        const int k = st.depth() - SMD;
        if (k >= LMD) return PB_DONE_OVF;
        l_ref[k] = far_ref;
        if (!ANY) l_t0[k] = far_t0;
      }
      st.up();
      return neg ? r1 : r0;
    }
    if (h1) return r1;
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
    int leaf_ok = -1;  // LEAF_EXACT: the reference's test of this leaf's box (bvh.rs:393), evaluated
                       // lazily with the maxt the leaf was entered with — no hit, no test
    for (uint32_t i = 0; i < cnt; ++i) {
      if (COUNT) res.steps += 2u;  // a primitive test costs about two node steps
      const uint32_t pi = off + i;
      uint32_t pr = SPH ? __ldg(&sc.leaf_prim[pi]) : pi;
      bool hit;
      float t, b1, b2 = 0.f;
      float4 a, b, c;
      if (SPH && (pr & PB_LEAF_BIT)) {
        hit = sphere_hit(sc.spheres + (pr & ~PB_LEAF_BIT), o, d, mint, maxt, &t, &b1);
      } else {
        const float4* tp = sc.tris + 3ull * pr;
        a = ldg4(tp), b = ldg4(tp + 1), c = ldg4(tp + 2);
        hit = tri_hit(mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), mk3(c.x, c.y, c.z), o, d, mint,
                      maxt, &t, &b1, &b2);
      }
      if (LEAF_EXACT && hit && leaf_ok < 0) {
        // maxt is still the value the leaf was entered with: this is the first hit inside it
        float T0;
        if (SPH || MULTI) {  // the leaf's own box as the reference stores it
          const float4 lo = ldg4(sc.leaf_boxes + 2ull * off), hi = ldg4(sc.leaf_boxes + 2ull * off + 1);
          leaf_ok = slab_test_finite(lo.x, lo.y, lo.z, hi.x, hi.y, hi.z, rb.o, rb.inv, mint, maxt, &T0) ? 1 : 0;
        } else {  // single-triangle leaf: mesh.rs:197-204, the bounds of its three vertices
          leaf_ok = slab_test_finite(fminf(fminf(a.x, b.x), c.x), fminf(fminf(a.y, b.y), c.y),
                                     fminf(fminf(a.z, b.z), c.z), fmaxf(fmaxf(a.x, b.x), c.x),
                                     fmaxf(fmaxf(a.y, b.y), c.y), fmaxf(fmaxf(a.z, b.z), c.z), rb.o,
                                     rb.inv, mint, maxt, &T0)
                        ? 1
                        : 0;
        }
      }
      if (LEAF_EXACT && leaf_ok == 0) return false;  // the reference never enters this leaf
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
  if (ANY && MODE == 4) {
    // Postponed leaves (any-hit only: the answer does not depend on the visiting order).  A lane that
    // reaches a leaf PARKS it and goes on with its stack; it only stops when it reaches a second leaf
    // or runs out of nodes.  The warp therefore tests triangles when every lane holds a parked leaf
    // (or is finished) instead of whenever any single lane does: in the lock-step model of
    // scripts/simt_cost.py the leaf rounds of a config-3 shadow packet drop from 4.7 to 2.8 and run
    // with 14 instead of 8 lanes.
    constexpr uint32_t NONE = PB_DONE;
    uint32_t pend = NONE;
    for (;;) {
      for (;;) {
        while (!(cur & PB_LEAF_BIT)) cur = node_step(cur);
        if (cur >= PB_DONE_OVF || pend != NONE) break;
        pend = cur;
        cur = pop();
      }
      if (pend == NONE) break;
      if (leaf(pend)) return res;
      pend = NONE;
    }
    if (cur == PB_DONE_OVF) res.prim = PB_OVERFLOW;
    return res;
  }
  if (SHAPE == 0) {
    while (cur < PB_DONE_OVF) {
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
      if (cur >= PB_DONE_OVF) break;
      if (leaf(cur)) return res;
      cur = pop();
    }
  }
  if (cur == PB_DONE_OVF) res.prim = PB_OVERFLOW;
  return res;
}

// BOX 3 constants.  Reference value r = fl(fl(p - o) * inv); ours f = fl(p * inv - c), c = fl(o * inv).
// With u = 2^-24 and every plane |p| <= B (the scene bound on that axis):
//   |r - x| <= (2u + u^2) |x|,  |f - x| <= u |o inv| (1 + u) + u |x|,  x = (p - o) inv,  |x| <= (B + |o|) |inv|
// so |f - r| < 4.1 u (B + |o|) |inv|.  e = 16 u (B + |o|) |inv| + FLT_MIN also covers the rounding of
// c -/+ e itself and results in the subnormal range.  Returns false when the bound is not usable
// (overflow); the caller then takes an exact path.
PB_DEV bool ffma_constants(const DScene& sc, f3 o, f3 inv, RayBox* rb) {
  const float k = 9.5367431640625e-07f;  // 2^-20
  const float tiny = 1.17549435e-38f;
  const float mx = (sc.babs[0] + fabsf(o.x)) * fabsf(inv.x), my = (sc.babs[1] + fabsf(o.y)) * fabsf(inv.y),
              mz = (sc.babs[2] + fabsf(o.z)) * fabsf(inv.z);
  if (!(fmaxf(fmaxf(mx, my), mz) < 1.2676506e30f)) return false;  // 2^100 (also rejects NaN)
  const float ex = mx * k + tiny, ey = my * k + tiny, ez = mz * k + tiny;
  const float cx = o.x * inv.x, cy = o.y * inv.y, cz = o.z * inv.z;
  rb->ncn = mk3(-(cx + ex), -(cy + ey), -(cz + ez));
  rb->ncf = mk3(-(cx - ex), -(cy - ey), -(cz - ez));
  return true;
}

// Which kernels get the 27-loop per-axis dispatch (else 8 octant loops + the generic one for warps
// that mix signs): measured, the any-hit kernel gains from it, the closest-hit kernel loses
// (instruction-cache pressure, profiles/r02_notes.md).
#ifndef PB_TRI_STATE
#define PB_TRI_STATE(ANY) (ANY)
#endif
#ifdef PB_HOST_CHECK
static int pb_host_force_mixed = 0;  // bit a: treat axis a as sign-mixed (tests/devsrc only)
#endif
// BOX: the box test the kernel variant was built for (1, 2 or 3; see child_box).  Rays it cannot
// serve take an exact fall-back inside the same kernel: a zero / NaN direction component -> BOX 0
// (the reference's compare-and-swap, NaN-faithful).  BOX 2 / 3 are specialised per axis for the sign
// of 1/d when the whole warp agrees on it (AX: 0 / 1), an axis on which the warp mixes signs keeps
// the min / max form (AX: 2) — one of 27 loops, selected by a warp-uniform switch.
template <bool ANY, bool SPH, bool MULTI, int MODE, int BOX = 1, bool COUNT = false>
PB_DEV TraceResult trace_ray(const DScene& sc, f3 o, f3 d, float mint, float maxt, uint32_t* s_ref,
                             float* s_t0) {
  RayBox rb;
  rb.o = o;
  rb.inv = mk3(1.f / d.x, 1.f / d.y, 1.f / d.z);  // bvh.rs:382
  rb.ncn = rb.ncf = mk3(0.f, 0.f, 0.f);
  const f3 inv = rb.inv;
  const float big = fmaxf(fmaxf(fabsf(inv.x), fabsf(inv.y)), fabsf(inv.z));
  // (NaN-propagating test: a NaN or infinite component takes the exact-compare path)
  const bool finite = big < __int_as_float(0x7f800000) && inv.x == inv.x && inv.y == inv.y && inv.z == inv.z;
  if (!finite || !sc.boxes_finite) return traverse<ANY, SPH, MULTI, 0, 0, 0, COUNT>(sc, o, d, rb, mint, maxt, s_ref, s_t0);
  if (BOX >= 2 && sc.boxes_ordered) {
    bool ok = true;
    if (BOX == 3) ok = ffma_constants(sc, o, inv, &rb);
    const unsigned m = __activemask();
    if (__all_sync(m, ok)) {
      // per axis: 0 = every lane >= 0, 1 = every lane < 0, 2 = mixed
      const unsigned bx = __ballot_sync(m, inv.x < 0.0f), by = __ballot_sync(m, inv.y < 0.0f),
                     bz = __ballot_sync(m, inv.z < 0.0f);
      int sx = bx == 0u ? 0 : (bx == m ? 1 : 2), sy = by == 0u ? 0 : (by == m ? 1 : 2),
          sz = bz == 0u ? 0 : (bz == m ? 1 : 2);
#ifdef PB_HOST_CHECK  // one emulated lane never mixes signs: the harness forces the mixed forms
      if (pb_host_force_mixed & 1) sx = 2;
      if (pb_host_force_mixed & 2) sy = 2;
      if (pb_host_force_mixed & 4) sz = 2;
#endif
#define PB_AX(X, Y, Z) \
  case X + 3 * Y + 9 * Z: \
    return traverse<ANY, SPH, MULTI, BOX, MODE, X | (Y << 2) | (Z << 4)>(sc, o, d, rb, mint, maxt, s_ref, s_t0);
      if (PB_TRI_STATE(ANY)) {  // 27 loops: shadow rays towards a light overhead mix signs in most warps
        switch (sx + 3 * sy + 9 * sz) {
#define PB_AX3(Y, Z) PB_AX(0, Y, Z) PB_AX(1, Y, Z) PB_AX(2, Y, Z)
#define PB_AX9(Z) PB_AX3(0, Z) PB_AX3(1, Z) PB_AX3(2, Z)
          PB_AX9(0) PB_AX9(1) PB_AX9(2)
#undef PB_AX9
#undef PB_AX3
        }
      } else if (sx != 2 && sy != 2 && sz != 2) {  // 8 loops: camera rays share their octant
        switch (sx + 3 * sy + 9 * sz) {
          PB_AX(0, 0, 0) PB_AX(1, 0, 0) PB_AX(0, 1, 0) PB_AX(1, 1, 0) PB_AX(0, 0, 1) PB_AX(1, 0, 1) PB_AX(0, 1, 1) PB_AX(1, 1, 1)
        }
      }
#undef PB_AX
    }
  }
  return traverse<ANY, SPH, MULTI, 1, MODE, 0, COUNT>(sc, o, d, rb, mint, maxt, s_ref, s_t0);
}
