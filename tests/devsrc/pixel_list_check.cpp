// CPU check of pbh::build_pixel_list (csrc/host_logic.hpp: per-interval task tables, separable need
// masks, threaded tile-major fill) against the straightforward per-pixel construction it replaced, which
// lives on here as the test's reference.  Built and run by tests/test_host_mirror.py.
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../../pbrt_rust_b200/csrc/host_logic.hpp"

namespace {
struct RefList {
  int r0 = 0, r1 = 0;
  std::vector<pbh::PixelRec> list;
  std::vector<int32_t> index;
  std::vector<uint32_t> rows_ready, row_first, keys;
};
inline const char* reference_pixel_list(const pbrtb200_sampler& smp, const int32_t* rects_p, size_t n_rects, bool whole, float xw,
                                    float yw, RefList* out) {
  struct RectView {
    const int32_t* p;
    const int32_t& operator[](size_t i) const { return p[i]; }
  } rects{rects_p};
  const int32_t ext[4] = {smp.x_start, smp.x_end, smp.y_start, smp.y_end};
  const int sw = ext[1] - ext[0], sh = ext[3] - ext[2];
  // Only the sampler rows this call can need are touched (row index relative to ext[2]): a row band of
  // a multi-GPU frame costs its share of the list build, not the whole frame's.  (A HaltonSampler bins
  // candidates that land anywhere: it keeps every row.)
  int r0 = 0, r1 = sh;
  auto rect_rows = [&](const int32_t* q, int* qy0, int* qy1) {
    *qy0 = std::max((int)std::ceil(((float)q[1] - 0.5f) - yw) - 1, ext[2]);
    *qy1 = std::min((int)std::floor(((float)(q[3] - 1) + 0.5f) + yw) + 1, ext[3] - 1);
  };
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
  auto loc = [&](int yy, int xx) { return (size_t)(yy - r0) * (size_t)sw + (size_t)xx; };  // yy, xx relative to the extent
  std::vector<uint16_t> task_of(n_loc, 0xFFFF);
  std::vector<uint32_t> k_of(n_loc, 0);
  std::vector<uint32_t>& keys = out->keys;
  keys.assign(8 * (size_t)smp.num_tasks, 0u);
  for (int t = 0; t < smp.num_tasks; ++t) {
    pbh::task_key((uint64_t)t, &keys[8 * (size_t)t]);
    int32_t w[4];
    pbh::sampler_sub_window(ext, (uint64_t)t, (uint64_t)smp.num_tasks, w);
    if (w[0] == w[1] || w[2] == w[3]) continue;  // get_sub_sampler -> None
    if (w[0] < ext[0] || w[1] > ext[1] || w[2] < ext[2] || w[3] > ext[3] || w[1] < w[0] || w[3] < w[2])
      return "task window outside the sampler extent";
    const uint32_t tw = (uint32_t)(w[1] - w[0]);
    for (int y = std::max(w[2], ext[2] + r0); y < std::min(w[3], ext[2] + r1); ++y)
      for (int x = w[0]; x < w[1]; ++x) {
        const size_t e = loc(y - ext[2], x - ext[0]);
        task_of[e] = (uint16_t)t;
        k_of[e] = (uint32_t)(y - w[2]) * tw + (uint32_t)(x - w[0]);
      }
  }
  // which sampler pixels are needed
  std::vector<uint8_t> need(n_loc, whole ? 1 : 0);
  std::vector<uint8_t> owned(whole ? 0 : n_loc, 0);  // sampler pixel lies inside a rect of this call
  if (!whole) {
    for (size_t r = 0; r < n_rects; ++r) {
      const int32_t* q = &rects[4 * r];
      for (int y = std::max(q[1], ext[2] + r0); y < std::min(q[3], ext[2] + r1); ++y)
        for (int x = std::max(q[0], ext[0]); x < std::min(q[2], ext[1]); ++x)
          owned[loc(y - ext[2], x - ext[0])] = 1;
      int qx0 = (int)std::ceil(((float)q[0] - 0.5f) - xw) - 1, qx1 = (int)std::floor(((float)(q[2] - 1) + 0.5f) + xw) + 1;
      int qy0, qy1;
      rect_rows(q, &qy0, &qy1);
      qx0 = std::max(qx0, ext[0]);
      qx1 = std::min(qx1, ext[1] - 1);
      qy0 = std::max(qy0, ext[2] + r0);
      qy1 = std::min(qy1, ext[2] + r1 - 1);
      // Within that padded range keep exactly the sampler pixels k_film will accept for some pixel
      // of the rect: a sample of pixel p has image coordinate in [p, p + 1], and add_sample's
      // extent arithmetic is monotonic, so it can only reach [ceil((p-0.5)-w), floor((p+0.5)+w)]
      // (the same float expressions k_film evaluates).
      auto reaches = [](int p, float w, int lo, int hi) {  // can pixel column/row p reach [lo, hi]?
        const int a = pbh::sat_i32(std::ceil(((float)p - 0.5f) - w));
        const int b = pbh::sat_i32(std::floor((((float)p + 1.0f) - 0.5f) + w));
        return a <= hi && b >= lo;
      };
      for (int y = qy0; y <= qy1; ++y) {
        if (!reaches(y, yw, q[1], q[3] - 1)) continue;
        for (int x = qx0; x <= qx1; ++x)
          if (reaches(x, xw, q[0], q[2] - 1)) need[loc(y - ext[2], x - ext[0])] = 1;
      }
    }
  }
  std::vector<pbh::PixelRec>& list = out->list;
  list.clear();
  list.reserve(n_loc);
  std::vector<int32_t>& index = out->index;
  index.assign(n_loc, -1);
  const int TW = 8, TH = 4;
  for (int ty = (r0 / TH) * TH; ty < r1; ty += TH)
    for (int tx = 0; tx < sw; tx += TW)
      for (int yy = std::max(ty, r0); yy < std::min(ty + TH, r1); ++yy)
        for (int xx = tx; xx < std::min(tx + TW, sw); ++xx) {
          const size_t e = loc(yy, xx);
          if (!need[e] || task_of[e] == 0xFFFF) continue;
          pbh::PixelRec p;
          const int x = ext[0] + xx, y = ext[2] + yy;
          p.xy = (int32_t)(((uint32_t)(uint16_t)(int16_t)x) | ((uint32_t)(uint16_t)(int16_t)y << 16));
          p.k = k_of[e];
          p.task = task_of[e];
          if (!whole && !owned[e]) p.task |= 0x80000000u;  // PB_PIXEL_HALO_BIT (scene.cuh)
          index[e] = (int32_t)list.size();
          list.push_back(p);
        }
  if (list.empty()) return "no sampler pixel to evaluate";
  // rows_ready[r] = how many list pixels must be finished before every sample of sampler rows
  // 0..r exists (lets k_film run on the finished top of the image while later chunks render)
  out->rows_ready.assign((size_t)sh, 0u);
  for (int yy = 0; yy < sh; ++yy) {
    uint32_t m = yy ? out->rows_ready[(size_t)yy - 1] : 0u;
    if (yy >= r0 && yy < r1)
      for (int xx = 0; xx < sw; ++xx) {
        const int32_t li = index[loc(yy, xx)];
        if (li >= 0) m = std::max(m, (uint32_t)li + 1u);
      }
    out->rows_ready[(size_t)yy] = m;
  }
  // row_first[r] = first list pixel still needed once everything above sampler row r is filtered
  out->row_first.assign((size_t)sh + 1, (uint32_t)list.size());
  for (int yy = sh - 1; yy >= 0; --yy) {
    uint32_t m = out->row_first[(size_t)yy + 1];
    if (yy >= r0 && yy < r1)
      for (int xx = 0; xx < sw; ++xx) {
        const int32_t li = index[loc(yy, xx)];
        if (li >= 0) m = std::min(m, (uint32_t)li);
      }
    out->row_first[(size_t)yy] = m;
  }
  out->r0 = r0;
  out->r1 = r1;
  return nullptr;
}


int check(const pbrtb200_sampler& s, const std::vector<int32_t>& rects, bool whole, float xw, float yw, const char* what) {
  RefList a;
  pbh::PixelList b;
  const char* ea = reference_pixel_list(s, rects.data(), rects.size() / 4, whole, xw, yw, &a);
  const char* eb = pbh::build_pixel_list(s, rects.data(), rects.size() / 4, whole, xw, yw, &b);
  if ((ea == nullptr) != (eb == nullptr)) {
    std::printf("FAIL %s: reference says %s, builder says %s\n", what, ea ? ea : "ok", eb ? eb : "ok");
    return 1;
  }
  if (ea) return 0;
  bool ok = a.r0 == b.r0 && a.r1 == b.r1 && a.list.size() == b.list.size() && a.index.size() == b.index.size() &&
            a.rows_ready == b.rows_ready && a.row_first == b.row_first && a.keys == b.keys;
  for (size_t i = 0; ok && i < a.list.size(); ++i)
    ok = a.list[i].xy == b.list[i].xy && a.list[i].k == b.list[i].k && a.list[i].task == b.list[i].task;
  for (size_t i = 0; ok && i < a.index.size(); ++i) ok = a.index[i] == b.index[i];
  if (!ok) std::printf("FAIL %s: lists differ (n %zu vs %zu, rows %d..%d vs %d..%d)\n", what, a.list.size(), b.list.size(), a.r0, a.r1, b.r0, b.r1);
  return ok ? 0 : 1;
}
}  // namespace

int main() {
  int bad = 0, n = 0;
  std::mt19937 rng(12345);
  auto U = [&](int lo, int hi) { return lo + (int)(rng() % (unsigned)(hi - lo + 1)); };
  for (int it = 0; it < 400; ++it) {
    pbrtb200_sampler s{};
    s.kind = (it % 9 == 8) ? PBRTB200_SAMPLER_HALTON : PBRTB200_SAMPLER_STRATIFIED;
    const int big = it % 40 == 0;
    const int w = big ? U(600, 2000) : U(1, 90), h = big ? U(300, 1100) : U(1, 70);
    const float widths[4] = {0.5f, 1.0f, 2.0f, 3.3f};
    const float xw = widths[U(0, 3)], yw = widths[U(0, 3)];
    // film pixels [fx0, fx0 + w) x [fy0, fy0 + h); sampler extent = the film's sample extent (film.rs:271-289)
    const int fx0 = U(-3, 5), fy0 = U(-3, 5);
    s.x_start = (int)std::floor((float)fx0 + 0.5f - xw);
    s.x_end = (int)std::ceil((float)(fx0 + w) - 0.5f + xw);
    s.y_start = (int)std::floor((float)fy0 + 0.5f - yw);
    s.y_end = (int)std::ceil((float)(fy0 + h) - 0.5f + yw);
    s.xs = s.ys = 2;
    s.jitter = 1;
    const int tasks[6] = {1, 2, 3, 13, 40, 128};
    s.num_tasks = tasks[U(0, 5)];
    char what[256];
    std::snprintf(what, sizeof what, "it %d film %dx%d@(%d,%d) filter %.1f/%.1f tasks %d kind %d", it, w, h, fx0, fy0, xw, yw, s.num_tasks, s.kind);
    bad += check(s, {fx0, fy0, fx0 + w, fy0 + h}, true, xw, yw, what);
    ++n;
    const int nr[5] = {1, 1, 2, 4, 7};
    const int k = nr[U(0, 4)];
    std::vector<int32_t> rects;
    for (int r = 0; r < k; ++r) {
      const int x0 = U(fx0, fx0 + w - 1), y0 = U(fy0, fy0 + h - 1);
      const int x1 = (r == 0 && it % 3 == 0) ? fx0 + w : U(x0 + 1, fx0 + w), y1 = U(y0 + 1, fy0 + h);
      rects.insert(rects.end(), {(r == 0 && it % 3 == 0) ? fx0 : x0, y0, x1, y1});
    }
    bad += check(s, rects, false, xw, yw, what);
    ++n;
  }
  std::printf("%s: %d lists compared, %d differ\n", bad ? "FAILED" : "OK", n, bad);
  return bad ? 1 : 0;
}
