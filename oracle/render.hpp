// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// SamplerRenderer::render restated (sampler_renderer.rs:37-206) with the SURVEY §0.2 decisions
// D1 (preprocess = no-op), D2 (dispatch to WhittedIntegrator::li), D3 (no volumes), D4 (NaN ->
// error), D5 (use the returned sample count, clear per-pixel buffers), D6 (film buffer is the
// output), D12 (num_tasks as written), D13 (strict mode keeps the sampler/film tile misalignment).
//
// Modes:
//   strict  — per-task sub-sampler + sub-film + RNG::new(task), add_sub_film overwrite.
//   default — identical per-pixel sample values (the task's stream addressed by pixel offset), but
//             every film pixel receives every sample within the filter radius, accumulated in
//             global raster order (y, x, sample).  This is the GPU path's contract.
#pragma once
#include <atomic>
#include <thread>

#include "shading.hpp"

namespace orc {

struct RenderConfig {
  PerspectiveCamera camera;
  Film film;          // full film (pixels zeroed)
  SamplerDesc sampler;  // full-extent sampler
  uint32_t num_tasks = 1;
  int mode = 0;  // 0 default, 1 strict
  int n_threads = 1;
  bool count_traversal = false;
  bool primary_only = false;  // skip shading (config-2 style hit-id runs)
};

struct RenderStats {
  uint64_t camera_rays = 0, camera_hits = 0, shadow_rays = 0;
  uint64_t nodes_visited = 0, tris_tested = 0, spheres_tested = 0;           // primary
  uint64_t sh_nodes_visited = 0, sh_tris_tested = 0, sh_spheres_tested = 0;  // shadow (any-hit)
  uint64_t nan_samples = 0;
  void add(const RenderStats& o) {
    camera_rays += o.camera_rays;
    camera_hits += o.camera_hits;
    shadow_rays += o.shadow_rays;
    nodes_visited += o.nodes_visited;
    tris_tested += o.tris_tested;
    spheres_tested += o.spheres_tested;
    sh_nodes_visited += o.sh_nodes_visited;
    sh_tris_tested += o.sh_tris_tested;
    sh_spheres_tested += o.sh_spheres_tested;
    nan_samples += o.nan_samples;
  }
};

struct TaskWindow {
  int ext[4];  // sampler sub-window
  uint32_t key[8];
  bool empty;
};

inline std::vector<TaskWindow> task_windows(const SamplerDesc& sd, uint32_t num_tasks) {
  std::vector<TaskWindow> tw(num_tasks);
  for (uint32_t t = 0; t < num_tasks; ++t) {
    compute_sub_window(sd.ext, t, num_tasks, tw[t].ext);
    tw[t].empty = (tw[t].ext[0] == tw[t].ext[1]) || (tw[t].ext[2] == tw[t].ext[3]);
    seed_from_u64(t, tw[t].key);  // sampler_renderer.rs:74  RNG::new(task_idx)
  }
  return tw;
}

// One camera sample: generate ray, trace, shade (SamplerRenderer::li, :184-206).
inline RGB sample_radiance(const Scene& sc, const RenderConfig& cfg, const CameraSample& cs,
                           const float* light_u, RenderStats* st, uint32_t* hit_prim,
                           float* hit_t) {
  RayDifferential rd = cfg.camera.generate_ray_differential(cs);
  rd.scale_differentials(1.0f / std::sqrt((float)cfg.sampler.spp()));  // :96
  st->camera_rays++;
  Hit h;
  TraceCounters tc;
  bool found = sc.bvh.intersect(rd.ray, &h, cfg.count_traversal ? &tc : nullptr);
  st->nodes_visited += tc.nodes_visited;
  st->tris_tested += tc.tris_tested;
  st->spheres_tested += tc.spheres_tested;
  if (hit_prim) *hit_prim = found ? h.prim : 0xFFFFFFFFu;
  if (hit_t) *hit_t = found ? h.t : 0.f;
  RGB L(0.0f);
  if (found) {
    st->camera_hits++;
    if (!cfg.primary_only) {
      ShadeCounters sh;
      L = whitted_li(sc, rd, h, light_u, cfg.count_traversal ? &sh : nullptr);
      if (cfg.count_traversal) {
        st->shadow_rays += sh.shadow_rays;
        st->sh_nodes_visited += sh.shadow_trace.nodes_visited;
        st->sh_tris_tested += sh.shadow_trace.tris_tested;
        st->sh_spheres_tested += sh.shadow_trace.spheres_tested;
      }
    }
  }
  // miss: sum of light.le(ray) = 0 (light/mod.rs:50-52)
  if (L.has_nans()) st->nan_samples++;  // D4
  return L;
}

// ---- HaltonSampler renders (sampler kind 2) -------------------------------------------------------
// The accepted candidates of every task, binned by home pixel.  Layout of everything per-sample
// (hit ids, the GPU's wavefront buffers): [pixel raster over the full extent][slot], `cap` slots per
// pixel (cap = the largest count), a pixel's samples in (task, index) order — one task per pixel.
struct HaltonEntry {
  uint32_t pixel;  // raster index over the full sampler extent
  uint32_t slot;
  uint32_t task;
  CameraSample cs;
};
struct HaltonLayout {
  std::vector<HaltonEntry> entries;  // task-major, index order within a task
  std::vector<float> light_u;        // 2 * light_samples floats per entry
  std::vector<uint32_t> counts;      // per pixel
  uint32_t cap = 0;
};
inline HaltonLayout halton_layout(const SamplerDesc& sd, uint32_t num_tasks, int n_threads) {
  if (sd.light_samples > ORC_HALTON_MAX_LIGHT_PAIRS) throw std::runtime_error("halton: too many light samples");
  std::vector<TaskWindow> tw = task_windows(sd, num_tasks);
  const int full_w = sd.ext[1] - sd.ext[0];
  const size_t lf = 2 * (size_t)sd.light_samples;
  std::vector<std::vector<HaltonEntry>> per(num_tasks);
  std::vector<std::vector<float>> per_lu(num_tasks);
  std::atomic<uint32_t> next{0};
  auto worker = [&]() {
    for (;;) {
      uint32_t t = next.fetch_add(1);
      if (t >= num_tasks) break;
      if (tw[t].empty) continue;
      HaltonWindow w = halton_window(tw[t].ext, sd.spp());
      std::vector<float> lu(lf + 1);
      for (uint64_t i = 0; i < w.wanted; ++i) {
        HaltonEntry e;
        if (!halton_candidate(sd, w, i, &e.cs, lu.data())) continue;
        int px, py;
        halton_home_pixel(w, e.cs, &px, &py);
        e.pixel = (uint32_t)((size_t)(py - sd.ext[2]) * (size_t)full_w + (size_t)(px - sd.ext[0]));
        e.slot = 0;
        e.task = t;
        per[t].push_back(e);
        per_lu[t].insert(per_lu[t].end(), lu.begin(), lu.begin() + lf);
      }
    }
  };
  std::vector<std::thread> th;
  for (int i = 0; i < std::max(1, n_threads); ++i) th.emplace_back(worker);
  for (auto& t : th) t.join();
  HaltonLayout L;
  L.counts.assign((size_t)full_w * (size_t)(sd.ext[3] - sd.ext[2]), 0u);
  for (uint32_t t = 0; t < num_tasks; ++t) {
    for (HaltonEntry& e : per[t]) {
      e.slot = L.counts[e.pixel]++;
      L.cap = std::max(L.cap, e.slot + 1);
      L.entries.push_back(e);
    }
    L.light_u.insert(L.light_u.end(), per_lu[t].begin(), per_lu[t].end());
  }
  return L;
}

inline void render_halton(const Scene& sc, RenderConfig& cfg, RenderStats* stats, uint32_t* hit_ids, float* hit_ts) {
  const SamplerDesc& sd = cfg.sampler;
  const int nthreads = std::max(1, cfg.n_threads);
  HaltonLayout L = halton_layout(sd, cfg.num_tasks, nthreads);
  const size_t n = L.entries.size(), lf = 2 * (size_t)sd.light_samples;
  if (hit_ids)
    for (size_t k = 0; k < L.counts.size() * (size_t)L.cap; ++k) {
      hit_ids[k] = 0xFFFFFFFFu;
      if (hit_ts) hit_ts[k] = 0.f;
    }
  std::vector<RGB> rad(n);
  std::vector<RenderStats> tstats((size_t)nthreads);
  std::atomic<size_t> next{0};
  auto worker = [&](int tid) {
    for (;;) {
      size_t b = next.fetch_add(256);
      if (b >= n) break;
      for (size_t k = b; k < std::min(n, b + 256); ++k) {
        const HaltonEntry& e = L.entries[k];
        size_t o = (size_t)e.pixel * (size_t)L.cap + e.slot;
        rad[k] = sample_radiance(sc, cfg, e.cs, L.light_u.data() + lf * k, &tstats[(size_t)tid],
                                 hit_ids ? hit_ids + o : nullptr, hit_ts ? hit_ts + o : nullptr);
      }
    }
  };
  std::vector<std::thread> th;
  for (int i = 0; i < nthreads; ++i) th.emplace_back(worker, i);
  for (auto& t : th) t.join();
  for (auto& s : tstats) stats->add(s);
  if (cfg.mode == 1) {
    // strict: per-task sub-films, samples in generation order (sampler_renderer.rs:61-144)
    size_t k = 0;
    for (uint32_t t = 0; t < cfg.num_tasks; ++t) {
      if (k >= n || L.entries[k].task != t) continue;
      Film tf = cfg.film.sub_film(t, cfg.num_tasks);
      for (; k < n && L.entries[k].task == t; ++k) tf.add_sample(L.entries[k].cs, rad[k].c);
      cfg.film.add_sub_film(tf);
    }
    return;
  }
  // default: pixel raster order, a pixel's samples in slot order (the GPU path's contract)
  std::vector<size_t> order(L.counts.size() * (size_t)L.cap, (size_t)-1);
  for (size_t k = 0; k < n; ++k) order[(size_t)L.entries[k].pixel * L.cap + L.entries[k].slot] = k;
  for (size_t o = 0; o < order.size(); ++o)
    if (order[o] != (size_t)-1) cfg.film.add_sample(L.entries[order[o]].cs, rad[order[o]].c);
}

// hit_ids / hit_ts (optional): one entry per camera sample, laid out
// [(y - ext.y0) * width + (x - ext.x0)] * spp + i over the FULL sampler extent.
inline void render(const Scene& sc, RenderConfig& cfg, RenderStats* stats, uint32_t* hit_ids,
                   float* hit_ts) {
  const SamplerDesc& sd = cfg.sampler;
  if (sd.kind == 2) {
    render_halton(sc, cfg, stats, hit_ids, hit_ts);
    return;
  }
  const size_t spp = sd.spp();
  const size_t W = sd.words_per_pixel();
  std::vector<TaskWindow> tw = task_windows(sd, cfg.num_tasks);
  const int full_w = sd.ext[1] - sd.ext[0];
  const int lsp = sd.light_samples;
  int nthreads = std::max(1, cfg.n_threads);

  if (cfg.mode == 1) {
    // ---- strict: one job per task (sampler_renderer.rs:61-144) ----
    std::vector<Film> task_films(cfg.num_tasks);
    std::vector<RenderStats> tstats(cfg.num_tasks);
    std::atomic<uint32_t> next{0};
    auto worker = [&]() {
      for (;;) {
        uint32_t t = next.fetch_add(1);
        if (t >= cfg.num_tasks) break;
        if (tw[t].empty) continue;  // get_sub_sampler -> None -> return
        task_films[t] = cfg.film.sub_film(t, cfg.num_tasks);
        RNG rng = RNG::from_key(tw[t].key);
        std::vector<CameraSample> cs;
        std::vector<float> lu;
        for (int y = tw[t].ext[2]; y < tw[t].ext[3]; ++y)
          for (int x = tw[t].ext[0]; x < tw[t].ext[1]; ++x) {
            pixel_samples(sd, x, y, rng, cs, lu);
            size_t base = ((size_t)(y - sd.ext[2]) * (size_t)full_w + (size_t)(x - sd.ext[0])) * spp;
            for (size_t i = 0; i < spp; ++i) {
              RGB L = sample_radiance(sc, cfg, cs[i], lu.data() + 2 * (size_t)lsp * i, &tstats[t],
                                      hit_ids ? hit_ids + base + i : nullptr,
                                      hit_ts ? hit_ts + base + i : nullptr);
              task_films[t].add_sample(cs[i], L.c);
            }
          }
      }
    };
    std::vector<std::thread> th;
    for (int i = 0; i < nthreads; ++i) th.emplace_back(worker);
    for (auto& t : th) t.join();
    for (uint32_t t = 0; t < cfg.num_tasks; ++t) {
      if (tw[t].empty) continue;
      cfg.film.add_sub_film(task_films[t]);
      stats->add(tstats[t]);
    }
    return;
  }

  // ---- default: same sample values, global raster-order accumulation ----
  const size_t n_px = (size_t)full_w * (size_t)(sd.ext[3] - sd.ext[2]);
  std::vector<CameraSample> all_cs(n_px * spp);
  std::vector<RGB> all_L(n_px * spp);
  struct Job {
    uint32_t task;
    int y;
  };
  std::vector<Job> jobs;
  for (uint32_t t = 0; t < cfg.num_tasks; ++t)
    if (!tw[t].empty)
      for (int y = tw[t].ext[2]; y < tw[t].ext[3]; ++y) jobs.push_back({t, y});
  std::atomic<size_t> next{0};
  std::vector<RenderStats> tstats((size_t)nthreads);
  auto worker = [&](int tid) {
    std::vector<CameraSample> cs;
    std::vector<float> lu;
    for (;;) {
      size_t j = next.fetch_add(1);
      if (j >= jobs.size()) break;
      const TaskWindow& w = tw[jobs[j].task];
      int y = jobs[j].y;
      RNG rng = RNG::from_key(w.key);
      size_t tw_w = (size_t)(w.ext[1] - w.ext[0]);
      rng.seek((size_t)(y - w.ext[2]) * tw_w * W);
      for (int x = w.ext[0]; x < w.ext[1]; ++x) {
        pixel_samples(sd, x, y, rng, cs, lu);
        size_t base = ((size_t)(y - sd.ext[2]) * (size_t)full_w + (size_t)(x - sd.ext[0])) * spp;
        for (size_t i = 0; i < spp; ++i) {
          all_cs[base + i] = cs[i];
          all_L[base + i] = sample_radiance(sc, cfg, cs[i], lu.data() + 2 * (size_t)lsp * i,
                                            &tstats[(size_t)tid],
                                            hit_ids ? hit_ids + base + i : nullptr,
                                            hit_ts ? hit_ts + base + i : nullptr);
        }
      }
    }
  };
  std::vector<std::thread> th;
  for (int i = 0; i < nthreads; ++i) th.emplace_back(worker, i);
  for (auto& t : th) t.join();
  for (auto& s : tstats) stats->add(s);
  for (size_t k = 0; k < all_cs.size(); ++k) cfg.film.add_sample(all_cs[k], all_L[k].c);
}

}  // namespace orc
