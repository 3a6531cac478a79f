// Analysis tool (not product, not a test): a lock-step model of one warp of the traversal kernels over
// the product's own pair nodes, to count node / leaf rounds and active lanes per 32-ray packet, and to
// compare traversal strategies before spending GPU time on them.  Built by scripts/simt_cost.py with
// g++ against the device headers in PB_HOST_CHECK mode.
#define PB_HOST_CHECK 1
#include <cstdint>
#include <cstdio>
#include <vector>

#include "../../pbrt_rust_b200/csrc/host_logic.hpp"
#include "../../pbrt_rust_b200/csrc/trace_core.cuh"

namespace {
struct Lane {
  f3 o, d;
  RayBox rb;
  float mint, maxt;
  uint32_t cur;
  std::vector<std::pair<uint32_t, float>> st;
  uint32_t prim;
  float t;
  bool done;
};
struct Counts {
  double node_rounds = 0, node_lanes = 0, leaf_rounds = 0, leaf_lanes = 0, push_rounds = 0, pop_rounds = 0, pop_iters = 0,
         box_tests = 0, tri_tests = 0, packets = 0, rays = 0, iters = 0;
};
struct Tree {
  DScene sc;
  pbh::PairNodes pn;
  std::vector<uint32_t> parent;     // pair index -> parent pair index (root: ~0)
  std::vector<uint32_t> leaf_pair;  // prim offset -> pair index holding that leaf as a child
};

inline bool box_exact(const Lane& L, float ax, float ay, float az, float bx, float by, float bz, float* T0) {
  return slab_test_finite(ax, ay, az, bx, by, bz, L.o, L.rb.inv, L.mint, L.maxt, T0);
}
}  // namespace

extern "C" {
void* simt_tree(const pbrtb200_scene* s) {
  Tree* t = new Tree();
  if (pbh::build_pair_nodes(s->nodes, s->n_nodes, s->n_prims, &t->pn)) return nullptr;
  t->sc = DScene{};
  t->sc.nodes = reinterpret_cast<const float4*>(t->pn.pairs.data());
  t->sc.tris = reinterpret_cast<const float4*>(s->tris);
  t->sc.root_ref = t->pn.root_ref;
  for (int i = 0; i < 3; ++i) {
    t->sc.root_bmin[i] = t->pn.root_bmin[i];
    t->sc.root_bmax[i] = t->pn.root_bmax[i];
  }
  const size_t np = t->pn.pairs.size() / 4;
  t->parent.assign(np, 0xFFFFFFFFu);
  t->leaf_pair.assign(s->n_prims, 0xFFFFFFFFu);
  for (size_t i = 0; i < np; ++i) {
    uint32_t r[2];
    std::memcpy(&r[0], &t->pn.pairs[4 * i + 3].x, 4);
    std::memcpy(&r[1], &t->pn.pairs[4 * i + 3].y, 4);
    for (int c = 0; c < 2; ++c) {
      if (r[c] & PB_LEAF_BIT)
        t->leaf_pair[r[c] & PB_LEAF_OFF_MASK] = (uint32_t)i;
      else
        t->parent[r[c]] = (uint32_t)i;
    }
  }
  return t;
}

// mode 0: closest hit, while-while.  mode 1: any-hit, if-if unordered.  Rays are consumed 32 at a time.
// hits_out (optional): prim, t per ray.  c12: the Counts fields.
void simt_run(void* tree, const float* rays8, uint64_t n, int mode, uint32_t* prim_out, float* t_out, double* c12) {
  Tree& T = *static_cast<Tree*>(tree);
  const DScene& sc = T.sc;
  Counts C;
  std::vector<Lane> W(32);
  for (uint64_t base = 0; base < n; base += 32) {
    const int nl = (int)std::min<uint64_t>(32, n - base);
    for (int l = 0; l < nl; ++l) {
      Lane& L = W[l];
      const float* r = rays8 + 8 * (base + l);
      L.o = mk3(r[0], r[1], r[2]);
      L.d = mk3(r[4], r[5], r[6]);
      L.mint = r[3];
      L.maxt = r[7];
      L.rb.o = L.o;
      L.rb.inv = mk3(1.f / L.d.x, 1.f / L.d.y, 1.f / L.d.z);
      L.st.clear();
      L.prim = PBRTB200_MISS;
      L.t = 0.f;
      float T0;
      L.done = !box_exact(L, sc.root_bmin[0], sc.root_bmin[1], sc.root_bmin[2], sc.root_bmax[0], sc.root_bmax[1], sc.root_bmax[2], &T0);
      L.cur = L.done ? PB_DONE : sc.root_ref;
    }
    C.packets += 1;
    C.rays += nl;
    auto pop = [&](Lane& L, int* iters) {
      *iters = 0;
      while (!L.st.empty()) {
        auto e = L.st.back();
        L.st.pop_back();
        ++*iters;
        if (mode == 1 || !(e.second > L.maxt)) return e.first;
      }
      return (uint32_t)PB_DONE;
    };
    auto node_step = [&](Lane& L, bool* pushed, int* pop_it) {
      const pbh::F4* q = &T.pn.pairs[4ull * L.cur];
      float T00, T01;
      const bool h0 = box_exact(L, q[0].x, q[0].y, q[0].z, q[0].w, q[1].x, q[1].y, &T00);
      const bool h1 = box_exact(L, q[1].z, q[1].w, q[2].x, q[2].y, q[2].z, q[2].w, &T01);
      C.box_tests += 2;
      uint32_t r0, r1, axis;
      std::memcpy(&r0, &q[3].x, 4);
      std::memcpy(&r1, &q[3].y, 4);
      std::memcpy(&axis, &q[3].w, 4);
      *pushed = false;
      *pop_it = 0;
      if (h0 && h1) {
        bool neg = false;
        if (mode == 0) neg = axis == 0 ? L.rb.inv.x < 0 : (axis == 1 ? L.rb.inv.y < 0 : L.rb.inv.z < 0);
        L.st.push_back({neg ? r0 : r1, neg ? T00 : T01});
        *pushed = true;
        L.cur = neg ? r1 : r0;
      } else if (h0) {
        L.cur = r0;
      } else if (h1) {
        L.cur = r1;
      } else {
        L.cur = pop(L, pop_it);
      }
    };
    auto leaf = [&](Lane& L, int* pop_it) {
      const uint32_t off = L.cur & PB_LEAF_OFF_MASK;
      const uint32_t cnt = ((L.cur >> PB_LEAF_CNT_SHIFT) & 0xFu) + 1u;
      bool answered = false;
      for (uint32_t i = 0; i < cnt && !answered; ++i) {
        const float4* tp = sc.tris + 3ull * (off + i);
        float t, b1, b2;
        C.tri_tests += 1;
        if (tri_hit(mk3(tp[0].x, tp[0].y, tp[0].z), mk3(tp[1].x, tp[1].y, tp[1].z), mk3(tp[2].x, tp[2].y, tp[2].z), L.o, L.d,
                    L.mint, L.maxt, &t, &b1, &b2)) {
          L.maxt = t;
          L.prim = off + i;
          L.t = t;
          if (mode == 1) answered = true;
        }
      }
      *pop_it = 0;
      L.cur = answered ? (uint32_t)PB_DONE : pop(L, pop_it);
    };
    for (;;) {
      bool any_active = false;
      for (int l = 0; l < nl; ++l) any_active |= W[l].cur != PB_DONE;
      if (!any_active) break;
      C.iters += 1;
      if (mode == 0) {  // while-while: node rounds until every lane holds a leaf / is done
        for (;;) {
          int k = 0, mp = 0;
          bool anyp = false;
          for (int l = 0; l < nl; ++l)
            if (!(W[l].cur & PB_LEAF_BIT)) {
              bool p;
              int it;
              node_step(W[l], &p, &it);
              anyp |= p;
              mp = std::max(mp, it);
              ++k;
            }
          if (!k) break;
          C.node_rounds += 1;
          C.node_lanes += k;
          C.push_rounds += anyp;
          C.pop_rounds += mp > 0;
          C.pop_iters += mp;
        }
        int k = 0, mp = 0;
        for (int l = 0; l < nl; ++l)
          if (W[l].cur != PB_DONE) {
            int it;
            leaf(W[l], &it);
            mp = std::max(mp, it);
            ++k;
          }
        if (k) {
          C.leaf_rounds += 1;
          C.leaf_lanes += k;
          C.pop_iters += mp;
        }
      } else {  // if-if: each iteration a lane does one node step OR one leaf
        int kn = 0, kl = 0, mp = 0;
        bool anyp = false;
        for (int l = 0; l < nl; ++l) {
          Lane& L = W[l];
          if (L.cur == PB_DONE) continue;
          int it = 0;
          if (!(L.cur & PB_LEAF_BIT)) {
            bool p;
            node_step(L, &p, &it);
            anyp |= p;
            ++kn;
          } else {
            leaf(L, &it);
            ++kl;
          }
          mp = std::max(mp, it);
        }
        if (kn) {
          C.node_rounds += 1;
          C.node_lanes += kn;
          C.push_rounds += anyp;
        }
        if (kl) {
          C.leaf_rounds += 1;
          C.leaf_lanes += kl;
        }
        C.pop_rounds += mp > 0;
        C.pop_iters += mp;
      }
    }
    for (int l = 0; l < nl; ++l) {
      if (prim_out) prim_out[base + l] = W[l].prim;
      if (t_out) t_out[base + l] = W[l].t;
    }
  }
  const double v[12] = {C.node_rounds, C.node_lanes, C.leaf_rounds, C.leaf_lanes, C.push_rounds, C.pop_rounds,
                        C.pop_iters,   C.box_tests,  C.tri_tests,   C.packets,    C.rays,        C.iters};
  for (int i = 0; i < 12; ++i) c12[i] = v[i];
}

// Bottom-up any-hit from the primitive the ray starts on (`from_prim` per ray): walk the parent chain
// of that leaf; at every ancestor test only the SIBLING subtree (the ancestors themselves contain the
// origin: by the containment lemma their tests are implied supersets), descending into siblings whose
// box passes with an ordinary top-down any-hit.  Counts box tests / tri tests / dependent node loads.
void simt_bottom_up(void* tree, const float* rays8, const uint32_t* from_prim, uint64_t n, uint8_t* occ_out, double* c4) {
  Tree& T = *static_cast<Tree*>(tree);
  const DScene& sc = T.sc;
  double box = 0, tri = 0, loads = 0, fallback = 0;
  std::vector<uint32_t> st;
  for (uint64_t i = 0; i < n; ++i) {
    Lane L;
    const float* r = rays8 + 8 * i;
    L.o = mk3(r[0], r[1], r[2]);
    L.d = mk3(r[4], r[5], r[6]);
    L.mint = r[3];
    L.maxt = r[7];
    L.rb.inv = mk3(1.f / L.d.x, 1.f / L.d.y, 1.f / L.d.z);
    bool occ = false;
    auto test_leaf = [&](uint32_t ref) {
      const uint32_t off = ref & PB_LEAF_OFF_MASK, cnt = ((ref >> PB_LEAF_CNT_SHIFT) & 0xFu) + 1u;
      for (uint32_t k = 0; k < cnt && !occ; ++k) {
        const float4* tp = sc.tris + 3ull * (off + k);
        float t, b1, b2;
        tri += 1;
        if (tri_hit(mk3(tp[0].x, tp[0].y, tp[0].z), mk3(tp[1].x, tp[1].y, tp[1].z), mk3(tp[2].x, tp[2].y, tp[2].z), L.o, L.d,
                    L.mint, L.maxt, &t, &b1, &b2))
          occ = true;
      }
    };
    auto descend = [&](uint32_t ref) {  // ordinary any-hit below `ref` (box of ref already passed)
      st.clear();
      uint32_t cur = ref;
      for (;;) {
        if (cur & PB_LEAF_BIT) {
          test_leaf(cur);
          if (occ || st.empty()) return;
          cur = st.back();
          st.pop_back();
          continue;
        }
        const pbh::F4* q = &T.pn.pairs[4ull * cur];
        loads += 1;
        float T00, T01;
        const bool h0 = box_exact(L, q[0].x, q[0].y, q[0].z, q[0].w, q[1].x, q[1].y, &T00);
        const bool h1 = box_exact(L, q[1].z, q[1].w, q[2].x, q[2].y, q[2].z, q[2].w, &T01);
        box += 2;
        uint32_t r0, r1;
        std::memcpy(&r0, &q[3].x, 4);
        std::memcpy(&r1, &q[3].y, 4);
        if (h0 && h1) {
          st.push_back(r1);
          cur = r0;
        } else if (h0)
          cur = r0;
        else if (h1)
          cur = r1;
        else {
          if (st.empty()) return;
          cur = st.back();
          st.pop_back();
        }
      }
    };
    const uint32_t p0 = from_prim[i];
    uint32_t pair = p0 < T.leaf_pair.size() ? T.leaf_pair[p0] : 0xFFFFFFFFu;
    if (pair == 0xFFFFFFFFu) {
      fallback += 1;
      descend(sc.root_ref);
    } else {
      // the start leaf itself (its triangle may still be hit: mint excludes the origin), then siblings upward
      uint32_t child_ref_from = 0xFFFFFFFFu;  // which child of `pair` we came from (leaf first)
      bool first = true;
      while (pair != 0xFFFFFFFFu && !occ) {
        const pbh::F4* q = &T.pn.pairs[4ull * pair];
        loads += 1;
        uint32_t r0, r1;
        std::memcpy(&r0, &q[3].x, 4);
        std::memcpy(&r1, &q[3].y, 4);
        int from;  // index of the child we came from
        if (first) {
          from = ((r0 & PB_LEAF_BIT) && (r0 & PB_LEAF_OFF_MASK) == p0) ? 0 : 1;
          // the origin leaf: reference semantics need its own box + triangle test too
          float T0;
          const bool h = from == 0 ? box_exact(L, q[0].x, q[0].y, q[0].z, q[0].w, q[1].x, q[1].y, &T0)
                                   : box_exact(L, q[1].z, q[1].w, q[2].x, q[2].y, q[2].z, q[2].w, &T0);
          box += 1;
          if (h) test_leaf(from == 0 ? r0 : r1);
          first = false;
        } else {
          from = (r0 == child_ref_from) ? 0 : 1;
        }
        if (occ) break;
        const int sib = 1 - from;
        float T0;
        const bool h = sib == 0 ? box_exact(L, q[0].x, q[0].y, q[0].z, q[0].w, q[1].x, q[1].y, &T0)
                                : box_exact(L, q[1].z, q[1].w, q[2].x, q[2].y, q[2].z, q[2].w, &T0);
        box += 1;
        if (h) descend(sib == 0 ? r0 : r1);
        child_ref_from = pair;
        pair = T.parent[pair];
      }
    }
    if (occ_out) occ_out[i] = occ ? 1 : 0;
  }
  c4[0] = box;
  c4[1] = tri;
  c4[2] = loads;
  c4[3] = fallback;
}
// Any-hit with POSTPONED leaves: a lane that reaches a leaf parks it (`slots` pending leaves per lane)
// and goes on with its stack; the warp runs a leaf round when `threshold` lanes hold a parked leaf, when
// a lane must park a leaf and has no free slot, or when no lane has node work left.  Order-free, hence
// legal for any-hit queries.  c: node_rounds, node_lanes, leaf_rounds, leaf_lanes, packets, rays, extra
// node steps done by lanes whose parked leaf turned out to be a hit.
void simt_postponed(void* tree, const float* rays8, uint64_t n, int threshold, int slots, uint8_t* occ_out, double* c8) {
  Tree& T = *static_cast<Tree*>(tree);
  const DScene& sc = T.sc;
  double node_rounds = 0, node_lanes = 0, leaf_rounds = 0, leaf_lanes = 0, packets = 0, nrays = 0, wasted = 0, tri = 0;
  struct PL {
    Lane L;
    std::vector<uint32_t> st;
    uint32_t pend[4];
    int np;
    bool occ, done;
    uint32_t cur;
  };
  std::vector<PL> W(32);
  for (uint64_t base = 0; base < n; base += 32) {
    const int nl = (int)std::min<uint64_t>(32, n - base);
    for (int l = 0; l < nl; ++l) {
      PL& P = W[l];
      const float* r = rays8 + 8 * (base + l);
      P.L.o = mk3(r[0], r[1], r[2]);
      P.L.d = mk3(r[4], r[5], r[6]);
      P.L.mint = r[3];
      P.L.maxt = r[7];
      P.L.rb.inv = mk3(1.f / P.L.d.x, 1.f / P.L.d.y, 1.f / P.L.d.z);
      P.st.clear();
      P.np = 0;
      P.occ = false;
      float T0;
      P.done = !box_exact(P.L, sc.root_bmin[0], sc.root_bmin[1], sc.root_bmin[2], sc.root_bmax[0], sc.root_bmax[1], sc.root_bmax[2], &T0);
      P.cur = P.done ? PB_DONE : sc.root_ref;
    }
    packets += 1;
    nrays += nl;
    auto popn = [&](PL& P) {
      if (P.st.empty()) return (uint32_t)PB_DONE;
      uint32_t r = P.st.back();
      P.st.pop_back();
      return r;
    };
    for (;;) {
      // node phase: every lane with an inner node does one step; a lane holding a leaf parks it
      int kn = 0, parked = 0, must_flush = 0, any_work = 0;
      for (int l = 0; l < nl; ++l) {
        PL& P = W[l];
        if (P.occ) continue;
        if (P.cur != PB_DONE && (P.cur & PB_LEAF_BIT)) {
          if (P.np < slots) {
            P.pend[P.np++] = P.cur;
            P.cur = popn(P);
          } else {
            must_flush = 1;
          }
        }
      }
      for (int l = 0; l < nl; ++l) {
        PL& P = W[l];
        if (P.occ || P.cur == PB_DONE || (P.cur & PB_LEAF_BIT)) continue;
        const pbh::F4* q = &T.pn.pairs[4ull * P.cur];
        float T00, T01;
        const bool h0 = box_exact(P.L, q[0].x, q[0].y, q[0].z, q[0].w, q[1].x, q[1].y, &T00);
        const bool h1 = box_exact(P.L, q[1].z, q[1].w, q[2].x, q[2].y, q[2].z, q[2].w, &T01);
        uint32_t r0, r1;
        std::memcpy(&r0, &q[3].x, 4);
        std::memcpy(&r1, &q[3].y, 4);
        if (h0 && h1) {
          P.st.push_back(r1);
          P.cur = r0;
        } else if (h0)
          P.cur = r0;
        else if (h1)
          P.cur = r1;
        else
          P.cur = popn(P);
        ++kn;
      }
      if (kn) {
        node_rounds += 1;
        node_lanes += kn;
      }
      for (int l = 0; l < nl; ++l) {
        PL& P = W[l];
        if (P.occ) continue;
        if (P.np) ++parked;
        if (P.cur != PB_DONE && !(P.cur & PB_LEAF_BIT)) any_work = 1;
        if (P.cur != PB_DONE && (P.cur & PB_LEAF_BIT) && P.np < slots) any_work = 1;  // can still park
      }
      if (parked && (must_flush || parked >= threshold || !any_work)) {
        int k = 0;
        for (int l = 0; l < nl; ++l) {
          PL& P = W[l];
          if (P.occ || !P.np) continue;
          ++k;
          // one leaf round tests ONE parked leaf per lane (the code path is one triangle test)
          const uint32_t ref = P.pend[--P.np];
          const uint32_t off = ref & PB_LEAF_OFF_MASK, cnt = ((ref >> PB_LEAF_CNT_SHIFT) & 0xFu) + 1u;
          for (uint32_t i = 0; i < cnt && !P.occ; ++i) {
            const float4* tp = sc.tris + 3ull * (off + i);
            float t, b1, b2;
            tri += 1;
            if (tri_hit(mk3(tp[0].x, tp[0].y, tp[0].z), mk3(tp[1].x, tp[1].y, tp[1].z), mk3(tp[2].x, tp[2].y, tp[2].z), P.L.o,
                        P.L.d, P.L.mint, P.L.maxt, &t, &b1, &b2))
              P.occ = true;
          }
        }
        leaf_rounds += 1;
        leaf_lanes += k;
      }
      bool alive = false;
      for (int l = 0; l < nl; ++l) {
        PL& P = W[l];
        if (!P.occ && (P.cur != PB_DONE || P.np)) alive = true;
      }
      if (!alive) break;
    }
    for (int l = 0; l < nl; ++l)
      if (occ_out) occ_out[base + l] = W[l].occ ? 1 : 0;
  }
  c8[0] = node_rounds; c8[1] = node_lanes; c8[2] = leaf_rounds; c8[3] = leaf_lanes; c8[4] = packets; c8[5] = nrays; c8[6] = wasted; c8[7] = tri;
}
void simt_free(void* t) { delete static_cast<Tree*>(t); }
}
