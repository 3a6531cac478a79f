// pbrtb200_group_*: all the GPUs of one box behind one call (include/pbrtb200.h).
//
// Replaces the fan-out of SamplerRenderer::render (src/sampler_renderer.rs:147-182): there a
// scoped_threadpool of num_cpus workers pulls image tiles and merges sub-films; here one persistent
// host thread per device renders one row band through pbrtb200_render's own pipeline.  No collective
// and no second process: the devices never exchange anything but the rows they store into the
// caller's film (host memory: each over its own PCIe link; device memory: peer stores over NVLink).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pbrtb200.h"

namespace {
thread_local std::string g_group_create_err;

struct Worker {
  int device = 0;
  pbrtb200_ctx* ctx = nullptr;
  std::thread th;
  int rc = 0;
  pbrtb200_stats st{};
};

// Row bands of equal cost over [y0, y1): bounds[k] snapped to `quantum` rows.
void cut_bands(const std::vector<double>& cost, int y0, int n, int quantum, std::vector<int>* bounds) {
  const int h = (int)cost.size();
  std::vector<double> pre((size_t)h + 1, 0.0);
  for (int y = 0; y < h; ++y) pre[(size_t)y + 1] = pre[(size_t)y] + std::max(cost[(size_t)y], 1e-9);
  bounds->assign((size_t)n + 1, y0);
  (*bounds)[(size_t)n] = y0 + h;
  int y = 0;
  for (int k = 1; k < n; ++k) {
    const double want = pre[(size_t)h] * k / n;
    while (y < h && pre[(size_t)y + 1] < want) ++y;
    int b = (int)std::lround((double)y / quantum) * quantum;
    b = std::min(std::max(b, (*bounds)[(size_t)k - 1] - y0), h);
    (*bounds)[(size_t)k] = y0 + b;
  }
}
}  // namespace

// The band balancer of a view (pure host arithmetic; also reachable through pbrtb200_bands_*).
// First frame: bands of equal summed cost from a per-row estimate (the cost probe).  Later frames:
// each band's cost density is rescaled by the time its device really needed (damped, exponent 0.8)
// and the bands are cut again while the slowest device is more than 3 % above the mean.  Every move
// makes the devices rebuild their pixel lists (a few ms of host work — several frames' worth at 8
// GPUs), so: at most 20 moves per view; after the first 8 frames only when two frames in a row say so
// (a single frame's device times carry ~2 % of noise); and once the bands have been within 3 % they
// are SETTLED: only three frames in a row more than 6 % off (the cost really changed, e.g. host
// instead of device film) start a new round — 3 % is one 4-row quantum of a 1080p band at 8 GPUs.
struct pbrtb200_bands {
  std::vector<double> row_cost;  // per film row: cost density estimate (probe, then measured)
  std::vector<int> bounds;       // n + 1 film rows
  int y0 = 0, n = 1, quantum = 4;
  int frames = 0, moves = 0, over_streak = 0;
  bool settled = false;  // the bands have been within 3 % once

  void reset(const std::vector<double>& cost, int y0_, int n_) {
    row_cost = cost;
    y0 = y0_;
    n = n_;
    frames = 1;
    moves = over_streak = 0;
    settled = false;
    cut_bands(row_cost, y0, n, quantum, &bounds);
  }
  // device_ms[k]: the time band k needed last frame.  Returns true when the boundaries moved.
  bool update(const float* device_ms) {
    ++frames;
    if (n < 2 || moves >= 20) return false;
    double mean = 0, mx = 0;
    for (int k = 0; k < n; ++k) {
      mean += device_ms[k] / n;
      mx = std::max<double>(mx, device_ms[k]);
    }
    if (!(mean > 0)) return false;
    if (mx <= 1.03 * mean) settled = true;
    const bool over = mx > (settled ? 1.06 : 1.03) * mean;
    over_streak = over ? over_streak + 1 : 0;
    const int need = settled ? 3 : (frames <= 9 ? 1 : 2);
    if (!over || over_streak < need) return false;
    settled = false;
    over_streak = 0;
    double total_cost = 0;
    for (double c : row_cost) total_cost += c;
    for (int k = 0; k < n; ++k) {
      const int a = bounds[(size_t)k] - y0, b = bounds[(size_t)k + 1] - y0;
      double band_cost = 0;
      for (int y = a; y < b; ++y) band_cost += row_cost[(size_t)y];
      if (b <= a || band_cost <= 0 || total_cost <= 0) continue;
      // predicted share of the frame vs the share of the time the device really needed
      const double predicted = band_cost / total_cost, measured = device_ms[k] / (mean * n);
      const double scale = std::pow(measured / predicted, 0.8);
      for (int y = a; y < b; ++y) row_cost[(size_t)y] *= scale;
    }
    const std::vector<int> before = bounds;
    cut_bands(row_cost, y0, n, quantum, &bounds);
    if (bounds == before) return false;
    ++moves;
    return true;
  }
};

struct pbrtb200_group {
  std::vector<Worker> w;
  std::string err;
  // one job at a time, handed to every worker
  std::mutex mu;
  std::condition_variable cv_go, cv_done;
  uint64_t epoch = 0;
  int pending = 0;
  bool quit = false;
  std::function<void(int)> job;
  // band state of the current view
  struct View {
    pbrtb200_camera cam{};
    pbrtb200_sampler smp{};
    int32_t film_ext[4] = {0, 0, 0, 0};
    float xw = 0, yw = 0;
    bool valid = false;
  } view;
  pbrtb200_bands bal;             // row bands of the current view
  std::vector<float> device_ms;
  bool peers_enabled = false;
  // host film page-locked on the caller's request (pbrtb200_group_pin_host_film)
  void* reg_ptr = nullptr;
  size_t reg_bytes = 0;

  // Hand-off: workers and the caller first SPIN on the atomics for a short while (back-to-back frames
  // of an interactive / benchmark loop: a condition-variable wake-up costs 20-50 us of a ~1 ms frame),
  // then sleep on the condition variables.
  std::atomic<uint64_t> a_epoch{0};
  std::atomic<int> a_pending{0};
  static constexpr int kSpin = 20000;
  void run(const std::function<void(int)>& f) {
    {
      std::lock_guard<std::mutex> lk(mu);
      job = f;
      pending = (int)w.size();
      a_pending.store(pending, std::memory_order_release);
      ++epoch;
      a_epoch.store(epoch, std::memory_order_release);
    }
    cv_go.notify_all();
    for (int k = 0; k < kSpin * 50 && a_pending.load(std::memory_order_acquire) != 0; ++k) std::this_thread::yield();
    std::unique_lock<std::mutex> lk(mu);
    cv_done.wait(lk, [&] { return pending == 0; });
  }
  void loop(int i) {
    cudaSetDevice(w[(size_t)i].device);
    uint64_t seen = 0;
    for (;;) {
      std::function<void(int)> f;
      for (int k = 0; k < kSpin && a_epoch.load(std::memory_order_acquire) == seen; ++k) std::this_thread::yield();
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_go.wait(lk, [&] { return quit || epoch != seen; });
        if (quit) return;
        seen = epoch;
        f = job;
      }
      f(i);
      {
        std::lock_guard<std::mutex> lk(mu);
        --pending;
        a_pending.store(pending, std::memory_order_release);
        if (pending == 0) cv_done.notify_all();
      }
    }
  }
  int fail(int code, const std::string& what) {
    err = what;
    return code;
  }
  // first error over the workers -> group error
  int collect(const char* what) {
    for (size_t i = 0; i < w.size(); ++i)
      if (w[i].rc != PBRTB200_OK) {
        err = std::string(what) + " on device " + std::to_string(w[i].device) + ": " + pbrtb200_last_error(w[i].ctx);
        return w[i].rc;
      }
    return PBRTB200_OK;
  }
};

extern "C" {

const char* pbrtb200_group_last_error(const pbrtb200_group* g) { return g ? g->err.c_str() : g_group_create_err.c_str(); }
int pbrtb200_group_size(const pbrtb200_group* g) { return g ? (int)g->w.size() : 0; }
pbrtb200_ctx* pbrtb200_group_ctx(pbrtb200_group* g, int i) {
  return (g && i >= 0 && i < (int)g->w.size()) ? g->w[(size_t)i].ctx : nullptr;
}

int pbrtb200_cut_bands(const float* row_cost, int n_rows, int y0, int n_bands, int32_t* bounds) {
  if (!row_cost || !bounds || n_rows < 1 || n_bands < 1) return PBRTB200_EINVAL;
  std::vector<double> c(row_cost, row_cost + n_rows);
  std::vector<int> b;
  cut_bands(c, y0, n_bands, 4, &b);
  for (int i = 0; i <= n_bands; ++i) bounds[i] = b[(size_t)i];
  return PBRTB200_OK;
}

pbrtb200_bands* pbrtb200_bands_new(const float* row_cost, int n_rows, int y0, int n_bands) {
  if (n_rows < 1 || n_bands < 1) return nullptr;
  std::vector<double> c((size_t)n_rows, 1.0);
  if (row_cost)
    for (int y = 0; y < n_rows; ++y) c[(size_t)y] = row_cost[y];
  pbrtb200_bands* b = new pbrtb200_bands();
  b->reset(c, y0, n_bands);
  return b;
}
void pbrtb200_bands_free(pbrtb200_bands* b) { delete b; }
int pbrtb200_bands_update(pbrtb200_bands* b, const float* device_ms) {
  if (!b || !device_ms) return PBRTB200_EINVAL;
  return b->update(device_ms) ? 1 : 0;
}
int pbrtb200_bands_get(const pbrtb200_bands* b, int32_t* bounds) {
  if (!b || !bounds) return PBRTB200_EINVAL;
  for (size_t i = 0; i < b->bounds.size(); ++i) bounds[i] = b->bounds[i];
  return PBRTB200_OK;
}

int pbrtb200_group_create(const int* devices, int n_devices, pbrtb200_group** out) {
  if (!out) return PBRTB200_EINVAL;
  *out = nullptr;
  int avail = 0;
  if (cudaGetDeviceCount(&avail) != cudaSuccess || avail == 0) {
    (void)cudaGetLastError();
    g_group_create_err = "no CUDA device";
    return PBRTB200_ENODEV;
  }
  if (n_devices < 1 || n_devices > avail) {
    g_group_create_err = "n_devices out of range (have " + std::to_string(avail) + ")";
    return PBRTB200_EINVAL;
  }
  pbrtb200_group* g = new pbrtb200_group();
  g->w.resize((size_t)n_devices);
  for (int i = 0; i < n_devices; ++i) {
    Worker& w = g->w[(size_t)i];
    w.device = devices ? devices[i] : i;
    for (int j = 0; j < i; ++j)
      if (g->w[(size_t)j].device == w.device) {
        g_group_create_err = "device listed twice";
        pbrtb200_group_destroy(g);
        return PBRTB200_EINVAL;
      }
    const int rc = pbrtb200_create(w.device, &w.ctx);
    if (rc != PBRTB200_OK) {
      g_group_create_err = std::string("device ") + std::to_string(w.device) + ": " + pbrtb200_last_error(nullptr);
      pbrtb200_group_destroy(g);
      return rc;
    }
  }
  // peer access towards the first device (device-resident films); absent peer access only disables
  // that output mode
  g->peers_enabled = true;
  for (int i = 1; i < n_devices; ++i) {
    int can = 0;
    cudaDeviceCanAccessPeer(&can, g->w[(size_t)i].device, g->w[0].device);
    if (!can) {
      g->peers_enabled = false;
      continue;
    }
    cudaSetDevice(g->w[(size_t)i].device);
    const cudaError_t e = cudaDeviceEnablePeerAccess(g->w[0].device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) g->peers_enabled = false;
    (void)cudaGetLastError();
  }
  for (int i = 0; i < n_devices; ++i) g->w[(size_t)i].th = std::thread([g, i] { g->loop(i); });
  *out = g;
  return PBRTB200_OK;
}

void pbrtb200_group_destroy(pbrtb200_group* g) {
  if (!g) return;
  {
    std::lock_guard<std::mutex> lk(g->mu);
    g->quit = true;
    g->cv_go.notify_all();
  }
  for (Worker& w : g->w)
    if (w.th.joinable()) w.th.join();
  if (g->reg_ptr) cudaHostUnregister(g->reg_ptr);
  for (Worker& w : g->w) {
    if (w.ctx) pbrtb200_destroy(w.ctx);
  }
  (void)cudaGetLastError();
  delete g;
}

int pbrtb200_group_upload_scene(pbrtb200_group* g, const pbrtb200_scene* scene) {
  if (!g) return PBRTB200_EINVAL;
  if (!scene) return g->fail(PBRTB200_EINVAL, "scene is NULL");
  g->view.valid = false;
  g->run([&](int i) { g->w[(size_t)i].rc = pbrtb200_upload_scene(g->w[(size_t)i].ctx, scene); });
  return g->collect("upload_scene");
}

// The group never page-locks a buffer on its own: a registration that outlives the caller's buffer
// would keep the OLD physical pages pinned, and a new buffer that malloc places at the same address
// would silently never receive its film.  The caller, who owns the buffer's lifetime, pins it.
int pbrtb200_group_pin_host_film(pbrtb200_group* g, float* xyzw, uint64_t bytes) {
  if (!g) return PBRTB200_EINVAL;
  if (!xyzw || bytes == 0) return g->fail(PBRTB200_EINVAL, "pin_host_film: NULL buffer or no bytes");
  pbrtb200_group_unpin_host_film(g);
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, xyzw) == cudaSuccess && at.type == cudaMemoryTypeHost) return PBRTB200_OK;  // already is
  (void)cudaGetLastError();
  if (cudaHostRegister(xyzw, (size_t)bytes, cudaHostRegisterPortable | cudaHostRegisterMapped) != cudaSuccess) {
    (void)cudaGetLastError();
    return g->fail(PBRTB200_ENOMEM, "pin_host_film: cudaHostRegister failed");
  }
  g->reg_ptr = xyzw;
  g->reg_bytes = (size_t)bytes;
  return PBRTB200_OK;
}
int pbrtb200_group_unpin_host_film(pbrtb200_group* g) {
  if (!g) return PBRTB200_EINVAL;
  if (g->reg_ptr) {
    cudaHostUnregister(g->reg_ptr);
    (void)cudaGetLastError();
  }
  g->reg_ptr = nullptr;
  g->reg_bytes = 0;
  return PBRTB200_OK;
}

int pbrtb200_group_device_stats(const pbrtb200_group* g, int i, pbrtb200_stats* out) {
  if (!g || !out || i < 0 || i >= (int)g->w.size()) return PBRTB200_EINVAL;
  *out = g->w[(size_t)i].st;
  return PBRTB200_OK;
}

int pbrtb200_group_bands(const pbrtb200_group* g, int32_t* bounds, float* device_ms) {
  if (!g || g->bal.bounds.size() != g->w.size() + 1) return PBRTB200_EINVAL;
  if (bounds)
    for (size_t i = 0; i < g->bal.bounds.size(); ++i) bounds[i] = g->bal.bounds[i];
  if (device_ms)
    for (size_t i = 0; i < g->w.size(); ++i) device_ms[i] = i < g->device_ms.size() ? g->device_ms[i] : 0.f;
  return PBRTB200_OK;
}

int pbrtb200_group_render(pbrtb200_group* g, const pbrtb200_camera* cam, const pbrtb200_sampler* smp,
                          const pbrtb200_film* film, const pbrtb200_integrator* integ, float* out_xyzw,
                          int out_is_device, pbrtb200_stats* stats) {
  if (!g) return PBRTB200_EINVAL;
  if (!cam || !smp || !film || !integ || !out_xyzw) return g->fail(PBRTB200_EINVAL, "NULL argument");
  if (film->x_pixel_count < 1 || film->y_pixel_count < 1) return g->fail(PBRTB200_EINVAL, "empty film");
  const int n = (int)g->w.size();
  if (stats) std::memset(stats, 0, sizeof *stats);
  if (out_is_device && n > 1 && !g->peers_enabled)
    return g->fail(PBRTB200_EINVAL, "device-resident group film needs peer access to the first device");
  const int W = film->x_pixel_count, H = film->y_pixel_count, y0 = film->y_pixel_start;

  // ---- bands -------------------------------------------------------------------------------------
  const int32_t ext[4] = {film->x_pixel_start, film->y_pixel_start, W, H};
  const bool same_view = g->view.valid && std::memcmp(&g->view.cam, cam, sizeof *cam) == 0 &&
                         std::memcmp(&g->view.smp, smp, sizeof *smp) == 0 && std::memcmp(g->view.film_ext, ext, sizeof ext) == 0 &&
                         g->view.xw == film->filter_xw && g->view.yw == film->filter_yw;
  if (!same_view) {
    std::vector<double> cost((size_t)H, 1.0);
    if (n > 1) {  // one-shot balance: the cost probe, on the first device
      std::vector<float> rc((size_t)H, 1.f);
      const int stride = std::max(1, std::min(W, H) / 256);
      if (pbrtb200_cost_profile(g->w[0].ctx, cam, film, stride, rc.data()) == PBRTB200_OK)
        for (int y = 0; y < H; ++y) cost[(size_t)y] = rc[(size_t)y];
    }
    g->bal.reset(cost, y0, n);
    g->view.cam = *cam;
    g->view.smp = *smp;
    std::memcpy(g->view.film_ext, ext, sizeof ext);
    g->view.xw = film->filter_xw;
    g->view.yw = film->filter_yw;
    g->view.valid = true;
  } else if (g->device_ms.size() == (size_t)n) {
    g->bal.update(g->device_ms.data());  // same view again: follow the measured device times
  }

  // ---- one band per device -----------------------------------------------------------------------
  g->run([&](int i) {
    Worker& w = g->w[(size_t)i];
    w.rc = PBRTB200_OK;
    std::memset(&w.st, 0, sizeof w.st);
    const int a = g->bal.bounds[(size_t)i], b = g->bal.bounds[(size_t)i + 1];
    if (b <= a) return;  // empty band: nothing to render, nothing to copy
    const int32_t rect[4] = {film->x_pixel_start, a, film->x_pixel_start + W, b};
    pbrtb200_tileset ts{rect, 1u, PBRTB200_TILES_KEEP_OTHERS};
    if (n == 1) {
      w.rc = pbrtb200_render(w.ctx, cam, smp, film, integ, nullptr, out_xyzw, out_is_device, &w.st);
      return;
    }
    if (out_is_device) {
      // the film lives on the first device: this device's film kernel stores its rows there
      w.rc = pbrtb200_render(w.ctx, cam, smp, film, integ, &ts, out_xyzw, 1, &w.st);
      return;
    }
    // Host film: pbrtb200_render with KEEP_OTHERS delivers exactly this band's rows into the caller's
    // buffer on this device's own stream — stored by k_film itself when the buffer is page-locked and
    // mapped (pbrtb200_group_pin_host_film, cudaHostAlloc), else staged in HBM and copied.
    w.rc = pbrtb200_render(w.ctx, cam, smp, film, integ, &ts, out_xyzw, 0, &w.st);
  });

  g->device_ms.assign((size_t)n, 0.f);
  bool nan = false;
  for (int i = 0; i < n; ++i) {
    const Worker& w = g->w[(size_t)i];
    g->device_ms[(size_t)i] = w.st.ms_total;
    if (w.rc == PBRTB200_ENAN) nan = true;
    if (stats) {
      stats->camera_rays += w.st.camera_rays;
      stats->camera_hits += w.st.camera_hits;
      stats->shadow_rays += w.st.shadow_rays;
      stats->kernel_launches += w.st.kernel_launches;
      stats->nan_samples += w.st.nan_samples;
      stats->stack_overflows += w.st.stack_overflows;
      if (w.st.ms_total >= stats->ms_total) {
        stats->ms_total = w.st.ms_total;
        stats->ms_raygen = w.st.ms_raygen;
        stats->ms_trace = w.st.ms_trace;
        stats->ms_shade = w.st.ms_shade;
        stats->ms_shadow = w.st.ms_shadow;
        stats->ms_film = w.st.ms_film;
      }
    }
  }
  for (int i = 0; i < n; ++i)
    if (g->w[(size_t)i].rc != PBRTB200_OK && g->w[(size_t)i].rc != PBRTB200_ENAN) return g->collect("render");
  if (nan) return g->fail(PBRTB200_ENAN, "Invalid radiance value!");
  return PBRTB200_OK;
}

}  // extern "C"
