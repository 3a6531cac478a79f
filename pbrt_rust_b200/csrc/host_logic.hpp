// Host-side logic the back end needs inside the library (plain C++, no CUDA): the per-task sampler
// sub-windows and RNG keys that SamplerRenderer derives before rendering, and the pixel work list.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/pbrtb200.h"
#include "leaf_ref.h"

namespace pbh {

// Rust `f32 as i32` / `as usize` (saturating; NaN -> 0)
inline int32_t sat_i32(float x) {
  if (x != x) return 0;
  if (x >= 2147483648.0f) return INT32_MAX;
  if (x <= -2147483648.0f) return INT32_MIN;
  return (int32_t)x;
}
inline uint64_t sat_usize(float x) {
  if (x != x || x <= 0.0f) return 0;
  if (x >= 18446744073709551616.0f) return UINT64_MAX;
  return (uint64_t)x;
}

// rand_core 0.6 SeedableRng::seed_from_u64 (PCG32 expansion) — the key of
// StdRng::seed_from_u64(task_idx), src/rng.rs:11-13.
inline void task_key(uint64_t task_idx, uint32_t key[8]) {
  uint64_t st = task_idx;
  for (int i = 0; i < 8; ++i) {
    st = st * 6364136223846793005ull + 11634580027462260723ull;
    const uint32_t xs = (uint32_t)(((st >> 18) ^ st) >> 27);
    const uint32_t rot = (uint32_t)(st >> 59);
    key[i] = (xs >> rot) | (xs << ((32u - rot) & 31u));
  }
}

// src/utils/mod.rs:171-205
inline void crop_window(uint64_t num, uint64_t count, float aspect, float w[4]) {
  auto split = [](uint64_t cnt, uint64_t asp, uint64_t* nx, uint64_t* ny) {
    uint64_t x = 1, y = cnt;
    while ((y % 2) == 0 && 2 * asp * x < y) {
      y /= 2;
      x *= 2;
    }
    *nx = x;
    *ny = y;
  };
  uint64_t nx, ny;
  if (aspect < 1.0f) {
    split(count, sat_usize(1.0f / aspect), &nx, &ny);
  } else {
    uint64_t a, b;
    split(count, sat_usize(aspect), &a, &b);
    nx = b;
    ny = a;
  }
  const uint64_t xo = num % nx, yo = num / nx;
  w[0] = (float)xo / (float)nx;
  w[1] = (float)(xo + 1) / (float)nx;
  w[2] = (float)yo / (float)ny;
  w[3] = (float)(yo + 1) / (float)ny;
}

// src/sampler/base.rs:29-48
inline void sampler_sub_window(const int32_t ext[4], uint64_t num, uint64_t count, int32_t out[4]) {
  const uint64_t dx = (uint64_t)(ext[1] - ext[0]), dy = (uint64_t)(ext[3] - ext[2]);
  const float aspect = (float)dx / (float)dy;
  float t[4];
  crop_window(num, count, aspect, t);
  const float psx = (float)ext[0], pex = (float)ext[1], psy = (float)ext[2], pey = (float)ext[3];
  auto lerp = [](float a, float b, float u) { return a * (1.0f - u) + b * u; };
  out[0] = sat_i32(lerp(psx, pex, t[0]));
  out[1] = sat_i32(lerp(psx, pex, t[1]));
  out[2] = sat_i32(lerp(psy, pey, t[2]));
  out[3] = sat_i32(lerp(psy, pey, t[3]));
}

// ---- BVH: the reference's linear PackedBVHNode array -> 64-byte pair nodes (layout: scene.cuh) ----
struct F4 {  // layout of CUDA's float4
  alignas(16) float x;
  float y, z, w;
};
struct PairNodes {
  std::vector<F4> pairs;             // 4 per inner node
  std::vector<uint16_t> leaf_count;  // per primitive offset: primitives of the leaf starting there
  std::vector<F4> leaf_boxes;        // 2 per primitive offset: box of the leaf starting there (min, max)
  float babs[3] = {0.f, 0.f, 0.f};   // max |coordinate| over all node boxes
  bool boxes_finite = true;          // no inf / NaN box coordinate
  bool boxes_ordered = true;         // min <= max on every axis of every box
  bool multi = false;                // some leaf holds more than one primitive
  bool big_leaf = false;             // some leaf holds >= 16 (count does not fit the inline field)
  uint32_t root_ref = 0;
  float root_bmin[3] = {0.f, 0.f, 0.f}, root_bmax[3] = {0.f, 0.f, 0.f};
};
// Returns nullptr on success, else what is wrong with the node array.
inline const char* build_pair_nodes(const pbrtb200_node32* nodes, uint32_t nn, uint32_t n_prims, PairNodes* out) {
  std::vector<uint32_t> pair_index(nn, 0xFFFFFFFFu);
  uint32_t n_inner = 0;
  for (uint32_t i = 0; i < nn; ++i)
    if (!nodes[i].is_leaf) pair_index[i] = n_inner++;
  out->leaf_count.assign(n_prims, 0);
  out->leaf_boxes.assign(2ull * n_prims, F4{0.f, 0.f, 0.f, 0.f});
  out->multi = out->big_leaf = false;
  out->boxes_finite = out->boxes_ordered = true;
  out->babs[0] = out->babs[1] = out->babs[2] = 0.f;
  for (uint32_t i = 0; i < nn; ++i)
    for (int a = 0; a < 3; ++a) {
      const float lo = nodes[i].bmin[a], hi = nodes[i].bmax[a];
      if (!std::isfinite(lo) || !std::isfinite(hi)) out->boxes_finite = false;
      if (!(lo <= hi)) out->boxes_ordered = false;
      out->babs[a] = std::max(out->babs[a], std::max(std::fabs(lo), std::fabs(hi)));
    }
  uint64_t covered = 0;
  auto leaf_ref = [](const pbrtb200_node32& nd) {
    return PB_LEAF_BIT | ((std::min<uint32_t>(nd.count, 16u) - 1u) << PB_LEAF_CNT_SHIFT) | nd.offset;
  };
  auto child_ref = [&](uint32_t c, uint32_t* ref) -> bool {
    if (c >= nn) return false;
    const pbrtb200_node32& nd = nodes[c];
    if (nd.is_leaf) {
      if (nd.count == 0 || (uint64_t)nd.offset + nd.count > n_prims) return false;
      *ref = leaf_ref(nd);
    } else {
      *ref = pair_index[c];
    }
    return true;
  };
  out->pairs.assign(4ull * n_inner, F4{0.f, 0.f, 0.f, 0.f});
  for (uint32_t i = 0; i < nn; ++i) {
    const pbrtb200_node32& nd = nodes[i];
    if (nd.is_leaf) {
      if (nd.count == 0 || (uint64_t)nd.offset + nd.count > n_prims)
        return "leaf node range outside the primitive list";
      out->leaf_count[nd.offset] = nd.count;
      out->leaf_boxes[2ull * nd.offset] = F4{nd.bmin[0], nd.bmin[1], nd.bmin[2], 0.f};
      out->leaf_boxes[2ull * nd.offset + 1] = F4{nd.bmax[0], nd.bmax[1], nd.bmax[2], 0.f};
      if (nd.count > 1) out->multi = true;
      if (nd.count >= 16) out->big_leaf = true;
      covered += nd.count;
      continue;
    }
    if (nd.axis > 2) return "inner node axis > 2";
    uint32_t r0, r1;
    if (i + 1 >= nn || nd.offset <= i + 1 || !child_ref(i + 1, &r0) || !child_ref(nd.offset, &r1))
      return "inner node child index invalid";
    const pbrtb200_node32 &c0 = nodes[i + 1], &c1 = nodes[nd.offset];
    F4* q = &out->pairs[4ull * pair_index[i]];
    q[0] = F4{c0.bmin[0], c0.bmin[1], c0.bmin[2], c0.bmax[0]};
    q[1] = F4{c0.bmax[1], c0.bmax[2], c1.bmin[0], c1.bmin[1]};
    q[2] = F4{c1.bmin[2], c1.bmax[0], c1.bmax[1], c1.bmax[2]};
    F4 m{0.f, 0.f, 0.f, 0.f};
    std::memcpy(&m.x, &r0, 4);
    std::memcpy(&m.y, &r1, 4);
    const uint32_t ax = nd.axis, ax_bit = 1u << nd.axis;
    std::memcpy(&m.z, &ax_bit, 4);
    std::memcpy(&m.w, &ax, 4);
    q[3] = m;
  }
  if (covered != n_prims) return "leaves do not cover the primitive list exactly once";
  const pbrtb200_node32& root = nodes[0];
  out->root_ref = root.is_leaf ? leaf_ref(root) : 0u;
  for (int i = 0; i < 3; ++i) {
    out->root_bmin[i] = root.bmin[i];
    out->root_bmax[i] = root.bmax[i];
  }
  return nullptr;
}

}  // namespace pbh
