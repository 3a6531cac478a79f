// Host-side logic the back end needs inside the library (plain C++, no CUDA): the per-task sampler
// sub-windows and RNG keys that SamplerRenderer derives before rendering, and the pixel work list.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
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

// ---- the pixel work list of a render call ------------------------------------------------------------
// The sampler pixels whose samples can reach the film pixels of `rects` (whole: every sampler pixel), in
// 8x4-tile-major order, each with its task (SamplerRenderer's sub-window split, sampler_renderer.rs:41-44
// + utils/mod.rs:171-205) and its raster index inside that task's window, which fixes its offset in the
// task's RNG stream.  Pure host arithmetic (api.cu uploads the result; pbrtb200_work_list exposes it).
struct PixelRec {  // layout of the device's DPixel (scene.cuh)
  int32_t xy;     // x | y << 16
  uint32_t k;     // raster index inside the task window
  uint32_t task;  // task index; bit 31: halo pixel (not inside a rect of this call)
};
// A grow-only array that is NOT zero-filled when it grows (value-initialising 33 MB costs more than
// filling it: the fill runs on several threads, each touching its own pages first).
template <class T>
struct RawBuf {
  std::unique_ptr<T[]> p;
  size_t n = 0, cap = 0;
  void resize_uninit(size_t m) {
    if (m > cap) {
      p.reset(new T[m]);  // default-initialised: trivial T stays uninitialised
      cap = m;
    }
    n = m;
  }
  size_t size() const { return n; }
  bool empty() const { return n == 0; }
  T* data() { return p.get(); }
  const T* data() const { return p.get(); }
  T& operator[](size_t i) { return p[i]; }
  const T& operator[](size_t i) const { return p[i]; }
  const T* begin() const { return p.get(); }
  const T* end() const { return p.get() + n; }
};
struct PixelList {
  int r0 = 0, r1 = 0;               // sampler rows (relative to the extent) the list covers
  RawBuf<PixelRec> list;
  RawBuf<int32_t> index;            // (r1 - r0) x width: list position of a sampler pixel, -1 = not listed
  std::vector<uint32_t> rows_ready; // [sampler row]: list pixels that must be finished before rows 0..r are complete
  std::vector<uint32_t> row_first;  // [sampler row]: first list pixel still needed once everything above row r is filtered
  std::vector<uint32_t> keys;       // 8 words per task: the ChaCha key of StdRng::seed_from_u64(task)
};
// rects: n_rects x (x0, y0, x1, y1) film pixels (validated by the caller); xw, yw: filter half-widths.
// Returns nullptr on success, else what is wrong.
inline const char* build_pixel_list(const pbrtb200_sampler& smp, const int32_t* rects, size_t n_rects, bool whole, float xw,
                                    float yw, PixelList* out) {
  const int32_t ext[4] = {smp.x_start, smp.x_end, smp.y_start, smp.y_end};
  const int sw = ext[1] - ext[0], sh = ext[3] - ext[2];
  if (sw <= 0 || sh <= 0 || smp.num_tasks < 1 || smp.num_tasks > 0xFFFE) return "bad sampler extent or task count";
  // Only the sampler rows this call can need are touched (row index relative to ext[2]): a row band of
  // a multi-GPU frame costs its share of the list build, not the whole frame's.  (A HaltonSampler bins
  // candidates that land anywhere: it keeps every row.)
  auto rect_rows = [&](const int32_t* q, int* qy0, int* qy1) {
    *qy0 = std::max((int)std::ceil(((float)q[1] - 0.5f) - yw) - 1, ext[2]);
    *qy1 = std::min((int)std::floor(((float)(q[3] - 1) + 0.5f) + yw) + 1, ext[3] - 1);
  };
  int r0 = 0, r1 = sh;
  if (!whole && smp.kind != PBRTB200_SAMPLER_HALTON) {
    r0 = sh;
    r1 = 0;
    for (size_t r = 0; r < n_rects; ++r) {
      int qy0, qy1;
      rect_rows(&rects[4 * r], &qy0, &qy1);
      r0 = std::min(r0, qy0 - ext[2]);
      r1 = std::max(r1, qy1 - ext[2] + 1);
    }
    r0 = std::max(0, std::min(r0, sh));
    r1 = std::max(r0, std::min(r1, sh));
  }
  const int rows_n = r1 - r0;
  const size_t n_loc = (size_t)sw * (size_t)rows_n;

  // ---- tasks: keys and sub-windows.  A pixel belongs to the LAST task whose window holds it (the
  // windows of the reference's split do not overlap; this keeps the old painting order anyway).
  // Rows between two consecutive window edges see the same windows, so the task of a pixel is looked
  // up in a per-row-interval table of `sw` entries instead of a per-pixel array.
  const int nt = smp.num_tasks;
  std::vector<uint32_t>& keys = out->keys;
  keys.assign(8 * (size_t)nt, 0u);
  struct Win {
    int32_t x0, x1, y0, y1;
    uint32_t tw;
  };
  std::vector<Win> win((size_t)nt);
  std::vector<int> edges = {ext[2] + r0, ext[2] + r1};
  for (int t = 0; t < nt; ++t) {
    task_key((uint64_t)t, &keys[8 * (size_t)t]);
    int32_t w[4];
    sampler_sub_window(ext, (uint64_t)t, (uint64_t)nt, w);
    Win& W = win[(size_t)t];
    W.x0 = w[0]; W.x1 = w[1]; W.y0 = w[2]; W.y1 = w[3];
    W.tw = (uint32_t)(w[1] - w[0]);
    if (w[0] == w[1] || w[2] == w[3]) {  // get_sub_sampler -> None
      W.x0 = W.x1 = W.y0 = W.y1 = 0;
      continue;
    }
    if (w[0] < ext[0] || w[1] > ext[1] || w[2] < ext[2] || w[3] > ext[3] || w[1] < w[0] || w[3] < w[2])
      return "task window outside the sampler extent";
    edges.push_back(std::min(std::max(w[2], ext[2] + r0), ext[2] + r1));
    edges.push_back(std::min(std::max(w[3], ext[2] + r0), ext[2] + r1));
  }
  std::sort(edges.begin(), edges.end());
  edges.erase(std::unique(edges.begin(), edges.end()), edges.end());
  const size_t n_iv = edges.size() - 1;  // row intervals [edges[i], edges[i + 1])
  std::vector<uint16_t> task_tab(std::max<size_t>(1, n_iv) * (size_t)sw, 0xFFFF);
  std::vector<uint16_t> iv_of_row((size_t)rows_n, 0);
  for (size_t i = 0; i < n_iv; ++i) {
    uint16_t* tab = &task_tab[i * (size_t)sw];
    for (int t = 0; t < nt; ++t) {
      const Win& W = win[(size_t)t];
      if (W.y0 <= edges[i] && edges[i + 1] <= W.y1 && W.x1 > W.x0)
        std::fill(tab + (W.x0 - ext[0]), tab + (W.x1 - ext[0]), (uint16_t)t);
    }
    for (int y = edges[i]; y < edges[i + 1]; ++y) iv_of_row[(size_t)(y - ext[2] - r0)] = (uint16_t)i;
  }

  // ---- which sampler pixels are needed, which are owned (inside a rect of this call) -----------------
  // need = OR over rects of (row can reach the rect) AND (column can reach the rect): kept as per-rect
  // row / column masks for up to 4 rects (one band of a multi-GPU frame: 1), per pixel beyond that.
  // A sample of pixel p has image coordinate in [p, p + 1], and add_sample's extent arithmetic is
  // monotonic, so it can only reach [ceil((p-0.5)-w), floor((p+0.5)+w)] (the float expressions k_film
  // evaluates).
  auto reaches = [](int p, float w, int lo, int hi) {  // can pixel column/row p reach [lo, hi]?
    const int a = sat_i32(std::ceil(((float)p - 0.5f) - w));
    const int b = sat_i32(std::floor((((float)p + 1.0f) - 0.5f) + w));
    return a <= hi && b >= lo;
  };
  const bool masks = !whole && n_rects <= 4;
  const bool per_pixel = !whole && !masks;
  // bit r of row_need[y] / col_need[x]: row / column can reach rect r; same for *_own: lies inside it
  std::vector<uint8_t> row_need, col_need, row_own, col_own, need_px, own_px;
  if (masks) {
    row_need.assign((size_t)rows_n, 0); row_own.assign((size_t)rows_n, 0);
    col_need.assign((size_t)sw, 0); col_own.assign((size_t)sw, 0);
    for (size_t r = 0; r < n_rects; ++r) {
      const int32_t* q = &rects[4 * r];
      const uint8_t bit = (uint8_t)(1u << r);
      int qx0 = (int)std::ceil(((float)q[0] - 0.5f) - xw) - 1, qx1 = (int)std::floor(((float)(q[2] - 1) + 0.5f) + xw) + 1;
      int qy0, qy1;
      rect_rows(q, &qy0, &qy1);
      qx0 = std::max(qx0, ext[0]); qx1 = std::min(qx1, ext[1] - 1);
      qy0 = std::max(qy0, ext[2] + r0); qy1 = std::min(qy1, ext[2] + r1 - 1);
      for (int y = qy0; y <= qy1; ++y)
        if (reaches(y, yw, q[1], q[3] - 1)) row_need[(size_t)(y - ext[2] - r0)] |= bit;
      for (int x = qx0; x <= qx1; ++x)
        if (reaches(x, xw, q[0], q[2] - 1)) col_need[(size_t)(x - ext[0])] |= bit;
      for (int y = std::max(q[1], ext[2] + r0); y < std::min(q[3], ext[2] + r1); ++y) row_own[(size_t)(y - ext[2] - r0)] |= bit;
      for (int x = std::max(q[0], ext[0]); x < std::min(q[2], ext[1]); ++x) col_own[(size_t)(x - ext[0])] |= bit;
    }
  } else if (per_pixel) {
    auto loc = [&](int yy, int xx) { return (size_t)(yy - r0) * (size_t)sw + (size_t)xx; };
    need_px.assign(n_loc, 0);
    own_px.assign(n_loc, 0);
    for (size_t r = 0; r < n_rects; ++r) {
      const int32_t* q = &rects[4 * r];
      for (int y = std::max(q[1], ext[2] + r0); y < std::min(q[3], ext[2] + r1); ++y)
        for (int x = std::max(q[0], ext[0]); x < std::min(q[2], ext[1]); ++x) own_px[loc(y - ext[2], x - ext[0])] = 1;
      int qx0 = (int)std::ceil(((float)q[0] - 0.5f) - xw) - 1, qx1 = (int)std::floor(((float)(q[2] - 1) + 0.5f) + xw) + 1;
      int qy0, qy1;
      rect_rows(q, &qy0, &qy1);
      qx0 = std::max(qx0, ext[0]); qx1 = std::min(qx1, ext[1] - 1);
      qy0 = std::max(qy0, ext[2] + r0); qy1 = std::min(qy1, ext[2] + r1 - 1);
      for (int y = qy0; y <= qy1; ++y) {
        if (!reaches(y, yw, q[1], q[3] - 1)) continue;
        for (int x = qx0; x <= qx1; ++x)
          if (reaches(x, xw, q[0], q[2] - 1)) need_px[loc(y - ext[2], x - ext[0])] = 1;
      }
    }
  }
  // flags of sampler pixel (row yy relative to r0, column xx): bit 0 needed, bit 1 owned
  auto flags_of = [&](int yy, int xx) -> unsigned {
    if (whole) return 3u;
    if (masks) return ((row_need[(size_t)yy] & col_need[(size_t)xx]) ? 1u : 0u) | ((row_own[(size_t)yy] & col_own[(size_t)xx]) ? 2u : 0u);
    const size_t e = (size_t)yy * (size_t)sw + (size_t)xx;
    return (need_px[e] ? 1u : 0u) | (own_px[e] ? 2u : 0u);
  };

  // ---- the list, 8x4-tile-major: counted per tile row, then filled (both passes run on a few threads) -
  const int TW = 8, TH = 4;
  const int ty_first = (r0 / TH) * TH;
  const int n_trows = rows_n > 0 ? (r1 - ty_first + TH - 1) / TH : 0;
  std::vector<uint32_t> trow_off((size_t)n_trows + 1, 0u);
  RawBuf<int32_t>& index = out->index;
  index.resize_uninit(n_loc);
  unsigned n_threads = 1;
  if (n_loc >= (1u << 18)) n_threads = std::min(std::max(1u, std::thread::hardware_concurrency()), std::min(8u, (unsigned)n_trows / 8u + 1u));
  auto parallel_rows = [&](auto&& body) {  // body(first tile row, last tile row)
    if (n_threads <= 1) {
      body(0, n_trows);
      return;
    }
    std::vector<std::thread> th;
    for (unsigned i = 0; i < n_threads; ++i) {
      const int a = (int)((long long)n_trows * i / n_threads), b = (int)((long long)n_trows * (i + 1) / n_threads);
      th.emplace_back([&body, a, b] { body(a, b); });
    }
    for (auto& t : th) t.join();
  };
  // rows (relative to r0) of tile row tr
  auto trow_rows = [&](int tr, int* ya, int* yb) {
    *ya = std::max(ty_first + tr * TH, r0) - r0;
    *yb = std::min(ty_first + (tr + 1) * TH, r1) - r0;
  };
  parallel_rows([&](int ta, int tb) {
    for (int tr = ta; tr < tb; ++tr) {
      int ya, yb;
      trow_rows(tr, &ya, &yb);
      uint32_t cnt = 0;
      for (int yy = ya; yy < yb; ++yy) {
        const uint16_t* tab = &task_tab[(size_t)iv_of_row[(size_t)yy] * (size_t)sw];
        for (int xx = 0; xx < sw; ++xx) cnt += ((flags_of(yy, xx) & 1u) && tab[xx] != 0xFFFF) ? 1u : 0u;
      }
      trow_off[(size_t)tr + 1] = cnt;
    }
  });
  for (int tr = 0; tr < n_trows; ++tr) trow_off[(size_t)tr + 1] += trow_off[(size_t)tr];
  const size_t n_list = n_trows ? trow_off[(size_t)n_trows] : 0;
  if (n_list == 0) return "no sampler pixel to evaluate";
  RawBuf<PixelRec>& list = out->list;
  list.resize_uninit(n_list);
  parallel_rows([&](int ta, int tb) {
    for (int tr = ta; tr < tb; ++tr) {
      int ya, yb;
      trow_rows(tr, &ya, &yb);
      uint32_t n = trow_off[(size_t)tr];
      for (int tx = 0; tx < sw; tx += TW) {
        const int xe = std::min(tx + TW, sw);
        for (int yy = ya; yy < yb; ++yy) {
          const uint16_t* tab = &task_tab[(size_t)iv_of_row[(size_t)yy] * (size_t)sw];
          int32_t* irow = &index[(size_t)yy * (size_t)sw];
          const int y = ext[2] + r0 + yy;
          for (int xx = tx; xx < xe; ++xx) {
            const unsigned f = flags_of(yy, xx);
            const uint16_t t = tab[xx];
            if (!(f & 1u) || t == 0xFFFF) {
              irow[xx] = -1;
              continue;
            }
            const Win& W = win[t];
            const int x = ext[0] + xx;
            PixelRec p;
            p.xy = (int32_t)(((uint32_t)(uint16_t)(int16_t)x) | ((uint32_t)(uint16_t)(int16_t)y << 16));
            p.k = (uint32_t)(y - W.y0) * W.tw + (uint32_t)(x - W.x0);
            p.task = (f & 2u) ? (uint32_t)t : ((uint32_t)t | 0x80000000u);  // PB_PIXEL_HALO_BIT (scene.cuh)
            irow[xx] = (int32_t)n;
            list[n++] = p;
          }
        }
      }
    }
  });

  // rows_ready[r] = how many list pixels must be finished before every sample of sampler rows 0..r
  // exists (lets k_film run on the finished top of the image while later chunks render);
  // row_first[r] = first list pixel still needed once everything above sampler row r is filtered.
  // Along one sampler row the list position grows with x (tiles to the right come later), so the last /
  // first listed pixel of the row carries the row's maximum / minimum.
  out->rows_ready.assign((size_t)sh, 0u);
  out->row_first.assign((size_t)sh + 1, (uint32_t)n_list);
  std::vector<int32_t> row_last((size_t)rows_n, -1), row_1st((size_t)rows_n, -1);
  for (int yy = 0; yy < rows_n; ++yy) {
    const int32_t* irow = &index[(size_t)yy * (size_t)sw];
    int xx = sw - 1;
    while (xx >= 0 && irow[xx] < 0) --xx;
    if (xx < 0) continue;
    row_last[(size_t)yy] = irow[xx];
    xx = 0;
    while (irow[xx] < 0) ++xx;
    row_1st[(size_t)yy] = irow[xx];
  }
  for (int yy = 0; yy < sh; ++yy) {
    uint32_t m = yy ? out->rows_ready[(size_t)yy - 1] : 0u;
    if (yy >= r0 && yy < r1 && row_last[(size_t)(yy - r0)] >= 0) m = std::max(m, (uint32_t)row_last[(size_t)(yy - r0)] + 1u);
    out->rows_ready[(size_t)yy] = m;
  }
  for (int yy = sh - 1; yy >= 0; --yy) {
    uint32_t m = out->row_first[(size_t)yy + 1];
    if (yy >= r0 && yy < r1 && row_1st[(size_t)(yy - r0)] >= 0) m = std::min(m, (uint32_t)row_1st[(size_t)(yy - r0)]);
    out->row_first[(size_t)yy] = m;
  }
  out->r0 = r0;
  out->r1 = r1;
  return nullptr;
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
