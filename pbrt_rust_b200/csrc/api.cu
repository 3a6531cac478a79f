// C-ABI implementation of the B200 back end (include/pbrtb200.h).  Host orchestration of the
// wavefront pipeline:  raygen -> trace(closest) -> shade -> trace(any) -> resolve -> film.
// There is no CPU fallback anywhere in this file: every entry point needs a CUDA device.
#include <cuda_runtime.h>

#include <chrono>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "film.cuh"
#include "host_logic.hpp"
#include "raygen.cuh"
#include "halton.cuh"
#include "shade.cuh"
#include "trace.cuh"
#include "trace_launch.h"

static_assert(sizeof(pbrtb200_node32) == 32, "node32 layout");
static_assert(sizeof(pbrtb200_tri48) == 48, "tri48 layout");
static_assert(sizeof(pbrtb200_sphere80) == 80, "sphere80 layout");
static_assert(sizeof(pbrtb200_ray32) == 32, "ray32 layout");
static_assert(sizeof(pbrtb200_hit16) == 16, "hit16 layout");
static_assert(sizeof(DAreaTri) == 64, "DAreaTri layout");

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  // grow-only
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    release();
    size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      e = cudaMalloc(&p, bytes);
      want = bytes;
    }
    if (e == cudaSuccess) cap = want;
    return e;
  }
  template <class T>
  T* as() const {
    return reinterpret_cast<T*>(p);
  }
};

struct CtrlBlock {  // small device-side control words, zeroed per use
  unsigned long long counter;
  unsigned long long shadow_total;
  unsigned long long hit_total;
  uint32_t flags;
  uint32_t sq_count;
  uint32_t nan_count;
  uint32_t pad;
};

thread_local std::string g_create_err;  // pbrtb200_create failures, per calling thread

}  // namespace

struct pbrtb200_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;      // the stream all work is issued on
  cudaStream_t own_stream = nullptr;  // created by the ctx
  cudaStream_t copy_stream = nullptr; // film bands travel to the host while later chunks render
  std::vector<cudaEvent_t> band_events;
  pbh::PixelList pixel_list;          // host copy of the work list (storage reused when the tile set changes)
  std::vector<uint32_t> rows_ready;   // per sampler row: list pixels that must be done (prefix max)
  std::vector<uint32_t> row_first;    // per sampler row r: smallest list index of any pixel in rows >= r
  std::string err;
  CtrlBlock* h_ctrl = nullptr;  // page-locked mirror of the device control block (read back once per call)
  // scene
  bool has_scene = false, has_spheres = false, multi_leaf = false;
  bool shade_ext = false;  // the scene needs k_shade's general texture evaluator / bump mapping
  DScene sc{};
  std::vector<pbrtb200_light> h_lights;
  DevBuf d_nodes, d_tris, d_leaf_prim, d_leaf_count, d_leaf_boxes, d_spheres, d_sphere_o2w, d_meshes, d_tri_uv,
      d_tri_n, d_tri_s, d_materials, d_mat_flags, d_textures, d_lights, d_area_tris, d_mipmaps, d_texels;
  std::vector<void*> peer_films;  // pbrtb200_peer_film_create allocations (freed at destroy)
  // per-frame work buffers (grow-only)
  DevBuf d_halton_tasks, d_hcounts, d_hidx, d_hoffsets;  // HaltonSampler: task windows, per-pixel counts, candidate indices, sample offsets
  std::vector<uint32_t> h_hoffsets;                      // host copy of the offsets (n_list_pixels + 1)
  unsigned long long halton_candidates = 0;
  uint32_t halton_n_tasks = 0;
  // The binning of a HaltonSampler frame is a pure function of the pixel list (no RNG): it is kept
  // until the list is rebuilt, like the list itself.
  bool halton_valid = false;   // d_hcounts / d_hidx hold the binned candidate indices of the list
  uint32_t halton_cap = 0;
  uint64_t halton_total = 0;
  DevBuf d_pixels, d_pix_index, d_task_keys, d_img, d_lens, d_time, d_lightu, d_edge, d_rec, d_terms, d_hits,
      d_sq_rays, d_sq_slots, d_film, d_rects, d_rect_prefix, d_ctrl, d_rays_in, d_occ, d_out_a,
      d_out_b, d_out_c;
  // cached pixel work list
  struct ListKey {
    pbrtb200_sampler smp{};
    int32_t film_ext[4] = {0, 0, 0, 0};
    float xw = 0, yw = 0;
    std::vector<int32_t> rects;
    bool whole = true;
    bool valid = false;
  } list_key;
  uint64_t n_list_pixels = 0;
  uint32_t n_film_pixels = 0;
  uint32_t n_rects = 0;
  std::vector<cudaEvent_t> events;
};

#define CK(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                      \
      (void)cudaGetLastError();                                                           \
      return e_ == cudaErrorMemoryAllocation ? PBRTB200_ENOMEM : PBRTB200_ENODEV;         \
    }                                                                                     \
  } while (0)

#define FAIL(code, msg) \
  do {                  \
    ctx->err = (msg);   \
    return (code);      \
  } while (0)

namespace {

template <class T>
int upload(pbrtb200_ctx* ctx, DevBuf& buf, const T* src, size_t n) {
  if (n == 0) return 0;
  CK(buf.ensure(n * sizeof(T)));
  CK(cudaMemcpyAsync(buf.p, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}

// Validates texture `id` and everything below it; returns its nesting depth or -1.  A texture with
// children (checkerboard, scale, mix, dots) may sit at levels 0 .. PBRTB200_TEX_MAX_DEPTH-1.
int tex_depth(const pbrtb200_scene* s, int id, int depth) {
  if (id < 0 || (uint32_t)id >= s->n_textures) return -1;
  const pbrtb200_texture& t = s->textures[id];
  if (t.kind < 0 || t.kind > PBRTB200_TEX_KIND_MAX) return -1;
  if (t.kind == PBRTB200_TEX_CONSTANT) return 0;
  if (t.kind == PBRTB200_TEX_FBM || t.kind == PBRTB200_TEX_WRINKLED)
    return (t.map_kind == PBRTB200_MAP_IDENTITY3D && t.aa >= 0) ? 0 : -1;
  const bool mapped = t.kind != PBRTB200_TEX_SCALE && t.kind != PBRTB200_TEX_MIX;
  if (mapped && (t.map_kind < 0 || t.map_kind > PBRTB200_MAP_CYLINDRICAL)) return -1;
  if (t.kind == PBRTB200_TEX_IMAGE) return (t.tex1 >= 0 && (uint32_t)t.tex1 < s->n_mipmaps) ? 0 : -1;
  if (t.kind == PBRTB200_TEX_UV || t.kind == PBRTB200_TEX_BILERP) return 0;
  if (depth >= PBRTB200_TEX_MAX_DEPTH) return -1;
  int d = 0;
  const int kids[3] = {t.tex1, t.tex2, t.kind == PBRTB200_TEX_MIX ? t.tex3 : t.tex2};
  for (int k : kids) {
    const int c = tex_depth(s, k, depth + 1);
    if (c < 0) return -1;
    d = std::max(d, c);
  }
  return 1 + d;
}
// Does the scene need the general texture evaluator (k_shade<.., EXT = true>)?
bool scene_needs_ext(const pbrtb200_scene* s) {
  for (uint32_t i = 0; i < s->n_textures; ++i)
    if (s->textures[i].kind > PBRTB200_TEX_IMAGE ||
        (s->textures[i].kind != PBRTB200_TEX_CONSTANT && s->textures[i].map_kind > PBRTB200_MAP_PLANAR))
      return true;
  for (uint32_t i = 0; i < s->n_materials; ++i)
    if (s->materials[i].bump != 0) return true;
  return false;
}

// Box-test family of the traversal kernels (trace_core.cuh child_box): 2 = exact tests specialised
// per sign of 1/d (default), 3 = FFMA tests for inner nodes + the reference's exact test in the leaf
// phase, 1 = exact min / max tests.  All three produce the same hits bit for bit; PBRTB200_BOX
// selects one for A/B measurements.
int trace_box_family() {
  static const int box = [] {
    const char* v = std::getenv("PBRTB200_BOX");
    const int b = v && *v ? std::atoi(v) : 2;
    return (b >= 1 && b <= 3) ? b : 2;
  }();
  return box;
}

template <bool ANY, int SRC>
int launch_trace_t(pbrtb200_ctx* ctx, const DCamera& cam, const TraceArgs& a_in) {
  // 32-ray packets a warp claims per work-counter atomic (same-address atomics retire at ~0.66 ns)
  static const int batch = [] {
    const char* v = std::getenv("PBRTB200_TRACE_BATCH");
    const int b = v && *v ? std::atoi(v) : 1;
    return b >= 1 && b <= 64 ? b : 1;
  }();
  TraceArgs a = a_in;
  a.batch = (uint32_t)batch;
  TraceLaunchCfg cfg{ctx->stream, ctx->sm_count, ctx->has_spheres, ctx->multi_leaf};
  CK(pb_launch_trace(ANY, SRC, trace_box_family(), cfg, ctx->sc, cam, a));
  return 0;
}

CtrlBlock* ctrl(pbrtb200_ctx* ctx) { return ctx->d_ctrl.as<CtrlBlock>(); }

void fill_camera(const pbrtb200_camera* c, int spp, DCamera* out) {
  std::memcpy(out->r2c, c->raster_to_camera, 64);
  std::memcpy(out->c2w, c->camera_to_world, 64);
  for (int i = 0; i < 3; ++i) {
    out->dx[i] = c->dx_camera[i];
    out->dy[i] = c->dy_camera[i];
  }
  out->sopen = c->shutter_open;
  out->sclose = c->shutter_close;
  out->lens_radius = c->lens_radius;
  out->focal_distance = c->focal_distance;
  out->diff_scale = 1.0f / std::sqrt((float)spp);  // sampler_renderer.rs:96
}

int sampler_spp(const pbrtb200_sampler* s) {
  if (s->kind == PBRTB200_SAMPLER_STRATIFIED) return s->xs * s->ys;
  if (s->kind == PBRTB200_SAMPLER_HALTON) return s->xs;  // samples_per_pixel as given (halton.rs:18-29)
  int p = 1;
  while (p < s->xs) p <<= 1;  // lds.rs:18 next_power_of_two
  return p;
}

int check_sampler(pbrtb200_ctx* ctx, const pbrtb200_sampler* s) {
  if (!s) FAIL(PBRTB200_EINVAL, "sampler is NULL");
  if (s->kind != PBRTB200_SAMPLER_STRATIFIED && s->kind != PBRTB200_SAMPLER_LD && s->kind != PBRTB200_SAMPLER_HALTON)
    FAIL(PBRTB200_EINVAL, "unsupported sampler kind");
  if (s->kind == PBRTB200_SAMPLER_HALTON && ctx->sc.area_sample_pairs > PB_HALTON_MAX_LIGHT_PAIRS)
    FAIL(PBRTB200_EINVAL, "HaltonSampler: more than 16 area-light samples per camera sample");
  if (s->xs < 1 || (s->kind == PBRTB200_SAMPLER_STRATIFIED && s->ys < 1))
    FAIL(PBRTB200_EINVAL, "sampler needs >= 1 sample per pixel");
  if (s->x_end <= s->x_start || s->y_end <= s->y_start)
    FAIL(PBRTB200_EINVAL, "empty sampler extent");
  if (s->num_tasks < 1 || s->num_tasks > 4096) FAIL(PBRTB200_EINVAL, "num_tasks out of range");
  if (s->x_start < -32768 || s->y_start < -32768 || s->x_end > 32767 || s->y_end > 32767)
    FAIL(PBRTB200_EINVAL, "sampler extent exceeds 16-bit pixel coordinates");
  return 0;
}

void fill_sampler(pbrtb200_ctx* ctx, const pbrtb200_sampler* s, DSampler* out) {
  out->kind = s->kind;
  out->xs = s->xs;
  out->ys = s->kind == PBRTB200_SAMPLER_STRATIFIED ? s->ys : 1;
  out->jitter = s->jitter;
  out->spp = sampler_spp(s);
  const uint32_t n = (uint32_t)out->spp;
  // RNG words per pixel (SURVEY §8a A1 / A1'): stratified 5n floats (+4n shuffle words), LD
  // (5 + 6n) u64 draws; the light-sample floats follow (D11).
  out->cam_words = s->kind == PBRTB200_SAMPLER_STRATIFIED ? (s->jitter ? 9u * n : 4u * n)
                                                          : 2u * (5u + 6u * n);
  out->words_per_pixel = out->cam_words + 2u * n * ctx->sc.area_sample_pairs;
  out->sopen = s->shutter_open;
  out->sclose = s->shutter_close;
  out->task_keys = ctx->d_task_keys.as<uint32_t>();
}

// Builds (or reuses) the pixel work list: the sampler pixels whose samples can reach the film
// pixels of `rects`, in 8x4-tile-major order, each with its task id and in-window index.
int build_pixel_list(pbrtb200_ctx* ctx, const pbrtb200_sampler* smp, const pbrtb200_film* film,
                     const pbrtb200_tileset* tiles) {
  // tiles == NULL: the whole film.  (A tile set WITHOUT rects never gets here: pbrtb200_render
  // answers it before building a list.)
  if (tiles && (tiles->n_rects == 0 || !tiles->rects)) FAIL(PBRTB200_EINVAL, "tile set without rects");
  const bool whole = tiles == nullptr;
  std::vector<int32_t> rects;
  const int32_t fx0 = film ? film->x_pixel_start : 0, fy0 = film ? film->y_pixel_start : 0;
  const int32_t fx1 = film ? fx0 + film->x_pixel_count : 0, fy1 = film ? fy0 + film->y_pixel_count : 0;
  if (!whole) {
    rects.assign(tiles->rects, tiles->rects + 4 * (size_t)tiles->n_rects);
    for (uint32_t r = 0; r < tiles->n_rects; ++r) {
      const int32_t* q = &rects[4 * r];
      if (q[0] < fx0 || q[1] < fy0 || q[2] > fx1 || q[3] > fy1 || q[2] <= q[0] || q[3] <= q[1])
        FAIL(PBRTB200_EINVAL, "tile rect outside the film pixel extent or empty");
    }
  } else if (film) {
    rects = {fx0, fy0, fx1, fy1};
  }
  auto& key = ctx->list_key;
  const float xw = film ? film->filter_xw : 0.f, yw = film ? film->filter_yw : 0.f;
  if (key.valid && std::memcmp(&key.smp, smp, sizeof *smp) == 0 && key.whole == whole &&
      key.rects == rects && key.xw == xw && key.yw == yw && key.film_ext[0] == fx0 &&
      key.film_ext[1] == fy0 && key.film_ext[2] == fx1 && key.film_ext[3] == fy1)
    return 0;
  key.valid = false;
  ctx->halton_valid = false;

  const int32_t ext[4] = {smp->x_start, smp->x_end, smp->y_start, smp->y_end};
  const int sw = ext[1] - ext[0], sh = ext[3] - ext[2];
  pbh::PixelList& pl = ctx->pixel_list;  // kept: a later list (a moved band) reuses the storage
  if (const char* why = pbh::build_pixel_list(*smp, rects.data(), rects.size() / 4, whole, xw, yw, &pl)) FAIL(PBRTB200_EINVAL, why);
  static_assert(sizeof(pbh::PixelRec) == sizeof(DPixel), "PixelRec is DPixel's host twin");
  const pbh::RawBuf<pbh::PixelRec>& list = pl.list;
  const pbh::RawBuf<int32_t>& index = pl.index;
  const std::vector<uint32_t>& keys = pl.keys;
  const int r0 = pl.r0;
  const size_t n_loc = index.size();
  ctx->rows_ready = pl.rows_ready;
  ctx->row_first = pl.row_first;
  if (upload(ctx, ctx->d_pixels, reinterpret_cast<const DPixel*>(list.data()), list.size())) return PBRTB200_ENODEV;
  // the index keeps its full-extent layout on the device (k_film / the Halton binning address it by
  // sampler row); only the rows of this list are rewritten — no other row is read with this list
  CK(ctx->d_pix_index.ensure((size_t)sw * (size_t)sh * sizeof(int32_t)));
  if (n_loc)
    CK(cudaMemcpyAsync(ctx->d_pix_index.as<int32_t>() + (size_t)r0 * (size_t)sw, index.data(), n_loc * sizeof(int32_t),
                       cudaMemcpyHostToDevice, ctx->stream));
  if (upload(ctx, ctx->d_task_keys, keys.data(), keys.size())) return PBRTB200_ENODEV;
  std::vector<DHaltonTask> htasks;
  if (smp->kind == PBRTB200_SAMPLER_HALTON) {  // HaltonSampler::new per sub-window (halton.rs:18-29, 36-47)
    unsigned long long first = 0;
    for (int t = 0; t < smp->num_tasks; ++t) {
      int32_t w[4];
      pbh::sampler_sub_window(ext, (uint64_t)t, (uint64_t)smp->num_tasks, w);
      DHaltonTask ht{};
      ht.x0 = w[0]; ht.x1 = w[1]; ht.y0 = w[2]; ht.y1 = w[3];
      ht.first = first;
      if (w[0] != w[1] && w[2] != w[3]) {
        const int dx = w[1] - w[0], dy = w[3] - w[2], m = dx > dy ? dx : dy;
        ht.wanted = (unsigned long long)((long long)m * (long long)m) * (unsigned long long)smp->xs;
        ht.delta = std::fmax((float)dy, (float)dx);
        if (ht.wanted > 0xFFFFFFFFull) FAIL(PBRTB200_EINVAL, "HaltonSampler: a task has more than 2^32 candidates");
      }
      first += ht.wanted;
      htasks.push_back(ht);
    }
    ctx->halton_candidates = first;
    ctx->halton_n_tasks = (uint32_t)htasks.size();
    if (upload(ctx, ctx->d_halton_tasks, htasks.data(), htasks.size())) return PBRTB200_ENODEV;
  }
  std::vector<uint32_t> prefix(rects.size() / 4 + 1, 0);
  for (size_t r = 0; r < rects.size() / 4; ++r)
    prefix[r + 1] = prefix[r] + (uint32_t)((rects[4 * r + 2] - rects[4 * r]) * (rects[4 * r + 3] - rects[4 * r + 1]));
  if (!rects.empty()) {
    if (upload(ctx, ctx->d_rects, rects.data(), rects.size())) return PBRTB200_ENODEV;
    if (upload(ctx, ctx->d_rect_prefix, prefix.data(), prefix.size())) return PBRTB200_ENODEV;
  }
  CK(cudaStreamSynchronize(ctx->stream));  // host vectors die at scope exit
  ctx->n_list_pixels = list.size();
  ctx->n_film_pixels = prefix.back();
  ctx->n_rects = (uint32_t)(rects.size() / 4);
  key.smp = *smp;
  key.whole = whole;
  key.rects = rects;
  key.xw = xw;
  key.yw = yw;
  key.film_ext[0] = fx0;
  key.film_ext[1] = fy0;
  key.film_ext[2] = fx1;
  key.film_ext[3] = fy1;
  key.valid = true;
  return 0;
}

}  // namespace
extern "C" int pbrtb200_work_list(const pbrtb200_sampler* smp, const pbrtb200_film* film, const pbrtb200_tileset* tiles,
                                  uint32_t* n_pixels, int32_t* xy, uint32_t* k, uint32_t* task, int32_t* index) {
  if (!smp || !film || !n_pixels || smp->x_end <= smp->x_start || smp->y_end <= smp->y_start || smp->num_tasks < 1)
    return PBRTB200_EINVAL;
  if (tiles && (tiles->n_rects == 0 || !tiles->rects)) return PBRTB200_EINVAL;
  std::vector<int32_t> rects;
  if (tiles)
    rects.assign(tiles->rects, tiles->rects + 4 * (size_t)tiles->n_rects);
  else
    rects = {film->x_pixel_start, film->y_pixel_start, film->x_pixel_start + film->x_pixel_count,
             film->y_pixel_start + film->y_pixel_count};
  pbh::PixelList pl;
  if (pbh::build_pixel_list(*smp, rects.data(), rects.size() / 4, tiles == nullptr, film->filter_xw, film->filter_yw, &pl))
    return PBRTB200_EINVAL;
  const bool fill = xy || k || task;
  if (fill && *n_pixels < pl.list.size()) return PBRTB200_EINVAL;
  *n_pixels = (uint32_t)pl.list.size();
  for (size_t i = 0; fill && i < pl.list.size(); ++i) {
    if (xy) xy[i] = pl.list[i].xy;
    if (k) k[i] = pl.list[i].k;
    if (task) task[i] = pl.list[i].task;
  }
  if (index) {
    const size_t sw = (size_t)(smp->x_end - smp->x_start), sh = (size_t)(smp->y_end - smp->y_start);
    std::fill(index, index + sw * sh, -1);
    std::copy(pl.index.begin(), pl.index.end(), index + (size_t)pl.r0 * sw);
  }
  return PBRTB200_OK;
}
namespace {

struct StageTimer {
  pbrtb200_ctx* ctx;
  bool on;
  size_t used = 0;
  struct Span {
    size_t a, b;
    int stage;
  };
  std::vector<Span> spans;
  cudaEvent_t get() {
    if (used == ctx->events.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      ctx->events.push_back(e);
    }
    return ctx->events[used++];
  }
  size_t mark() {
    if (!on) return 0;
    cudaEvent_t e = get();
    cudaEventRecord(e, ctx->stream);
    return used - 1;
  }
  void span(size_t a, size_t b, int stage) {
    if (on) spans.push_back({a, b, stage});
  }
  void collect(float ms[6]) {
    for (int i = 0; i < 6; ++i) ms[i] = 0.f;
    if (!on) return;
    for (auto& s : spans) {
      float t = 0.f;
      cudaEventElapsedTime(&t, ctx->events[s.a], ctx->events[s.b]);
      ms[s.stage] += t;
    }
  }
};

// Samples per wavefront chunk: bounds the chunk-local buffers (hits, Le, contrib, shadow queue).
// 2^25: measured best of 2^21..2^26 once the shade stage stopped being atomic-bound (every launch
// of a persistent trace kernel costs ~40 us of ramp and tail; profiles/r01_notes.md).  render()
// halves it until the chunk-local buffers fit kChunkBudgetBytes (many light slots).
constexpr int kChunkLog2Default = 25;
constexpr uint64_t kChunkBudgetBytes = 4ull << 30;   // chunk-local buffers (hits, shadow queue, terms)
// Per-sample frame buffers larger than this go into the pixel ring (render()).  PBRTB200_FRAME_BUDGET_MB
// overrides it (read per call: the tests force the ring on small frames with it).
uint64_t frame_budget_bytes() {
  const char* v = std::getenv("PBRTB200_FRAME_BUDGET_MB");
  const long long mb = v && *v ? std::atoll(v) : 4096;
  return (uint64_t)std::max(1ll, mb) << 20;
}
uint64_t chunk_samples() {  // (read per call: the tests shrink it to run several chunks on small frames)
  const char* v = std::getenv("PBRTB200_CHUNK_LOG2");
  const int x = v && *v ? std::atoi(v) : kChunkLog2Default;
  return 1ull << ((x >= 10 && x <= 30) ? x : kChunkLog2Default);
}

}  // namespace

extern "C" {

const char* pbrtb200_last_error(const pbrtb200_ctx* ctx) {
  return ctx ? ctx->err.c_str() : g_create_err.c_str();
}

int pbrtb200_create(int device, pbrtb200_ctx** out) {
  if (!out) return PBRTB200_EINVAL;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    g_create_err = std::string("no CUDA device: ") + cudaGetErrorString(e);
    (void)cudaGetLastError();
    return PBRTB200_ENODEV;
  }
  if (device < 0 || device >= n) {
    g_create_err = "device index out of range";
    return PBRTB200_EINVAL;
  }
  pbrtb200_ctx* ctx = new pbrtb200_ctx();
  ctx->device = device;
  auto bail = [&](const char* what, cudaError_t err) {
    g_create_err = std::string(what) + ": " + cudaGetErrorString(err);
    (void)cudaGetLastError();
    delete ctx;
    return PBRTB200_ENODEV;
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail("cudaGetDeviceProperties", e);
  ctx->sm_count = prop.multiProcessorCount;
  if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess)
    return bail("cudaStreamCreate", e);
  ctx->stream = ctx->own_stream;
  if ((e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) != cudaSuccess)
    return bail("cudaStreamCreate", e);
  if ((e = ctx->d_ctrl.ensure(sizeof(CtrlBlock))) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaMallocHost(&ctx->h_ctrl, sizeof(CtrlBlock))) != cudaSuccess) return bail("cudaMallocHost", e);
  *out = ctx;
  return PBRTB200_OK;
}

void pbrtb200_destroy(pbrtb200_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (cudaEvent_t e : ctx->events) cudaEventDestroy(e);
  cudaStreamDestroy(ctx->own_stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  for (void* p : ctx->peer_films) cudaFree(p);
  if (ctx->h_ctrl) cudaFreeHost(ctx->h_ctrl);
  for (cudaEvent_t ev : ctx->band_events) cudaEventDestroy(ev);
  delete ctx;
}

int pbrtb200_set_stream(pbrtb200_ctx* ctx, void* cuda_stream) {
  if (!ctx) return PBRTB200_EINVAL;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
  return PBRTB200_OK;
}

int pbrtb200_upload_scene(pbrtb200_ctx* ctx, const pbrtb200_scene* s) {
  if (!ctx) return PBRTB200_EINVAL;
  if (!s) FAIL(PBRTB200_EINVAL, "scene is NULL");
  CK(cudaSetDevice(ctx->device));
  ctx->has_scene = false;
  ctx->list_key.valid = false;
  if (s->n_nodes == 0 || !s->nodes) FAIL(PBRTB200_EINVAL, "scene has no BVH nodes");
  if (s->n_prims == 0) FAIL(PBRTB200_EINVAL, "scene has no primitives");
  if (s->n_prims > PB_LEAF_OFF_MASK - 16u) FAIL(PBRTB200_EINVAL, "too many primitives (limit 134217711)");
  if (s->n_spheres && (!s->leaf_prim || !s->spheres || !s->sphere_o2w))
    FAIL(PBRTB200_EINVAL, "spheres need leaf_prim, spheres and sphere_o2w");
  if (!s->n_spheres && s->n_tris != s->n_prims && !s->leaf_prim)
    FAIL(PBRTB200_EINVAL, "n_tris != n_prims without leaf_prim");
  if (s->n_tris && !s->tris) FAIL(PBRTB200_EINVAL, "tris is NULL");
  if (s->n_tris && (!s->meshes || !s->n_meshes)) FAIL(PBRTB200_EINVAL, "triangles need meshes");

  // ---- BVH: reference linear nodes -> 64-byte pair nodes (host_logic.hpp) ----
  pbh::PairNodes pn;
  if (const char* why = pbh::build_pair_nodes(s->nodes, s->n_nodes, s->n_prims, &pn)) FAIL(PBRTB200_EINVAL, why);
  static_assert(sizeof(pbh::F4) == sizeof(float4), "pair-node quads are float4");
  const std::vector<pbh::F4>& pairs = pn.pairs;
  const std::vector<uint16_t>& leaf_count = pn.leaf_count;
  bool multi = pn.multi;
  const bool big_leaf = pn.big_leaf;
  // The FFMA kernels apply the reference's exact test of a leaf's box in the leaf phase.  For a leaf
  // that is one triangle the box is recomputed from the vertices (mesh.rs:197-204) — if the caller's
  // node array agrees (+0 == -0); otherwise, and for quadrics / multi-primitive leaves, the stored
  // boxes are read from memory (the MULTI kernel variants).
  if (!multi && !(s->n_spheres > 0 || s->leaf_prim) && s->tris) {
    for (uint32_t i = 0; i < s->n_nodes && !multi; ++i) {
      const pbrtb200_node32& nd = s->nodes[i];
      if (!nd.is_leaf) continue;
      const pbrtb200_tri48& t = s->tris[nd.offset];
      const float* v[3] = {t.p1, t.p2, t.p3};
      for (int a = 0; a < 3; ++a) {
        const float lo = std::fmin(std::fmin(v[0][a], v[1][a]), v[2][a]), hi = std::fmax(std::fmax(v[0][a], v[1][a]), v[2][a]);
        if (!(lo == nd.bmin[a]) || !(hi == nd.bmax[a])) multi = true;
      }
    }
  }

  // ---- validation of the shading tables ----
  for (uint32_t i = 0; i < s->n_materials; ++i) {
    const pbrtb200_material& m = s->materials[i];
    if (m.kind != PBRTB200_MAT_MATTE && m.kind != PBRTB200_MAT_PLASTIC)
      FAIL(PBRTB200_EINVAL, "unsupported material kind");
    if (tex_depth(s, m.kd, 0) < 0) FAIL(PBRTB200_EINVAL, "material kd texture invalid or nested too deep");
    if (m.kind == PBRTB200_MAT_MATTE && tex_depth(s, m.sigma, 0) < 0)
      FAIL(PBRTB200_EINVAL, "material sigma texture invalid");
    if (m.kind == PBRTB200_MAT_PLASTIC && (tex_depth(s, m.ks, 0) < 0 || tex_depth(s, m.roughness, 0) < 0))
      FAIL(PBRTB200_EINVAL, "material ks/roughness texture invalid");
    if (m.bump != 0 && tex_depth(s, m.bump - 1, 0) < 0) FAIL(PBRTB200_EINVAL, "material bump texture invalid");
  }
  for (uint32_t i = 0; i < s->n_mipmaps; ++i) {  // every level must lie inside the texel pool
    const pbrtb200_mipmap& m = s->mipmaps[i];
    if (m.width == 0 || m.height == 0 || (m.width & (m.width - 1)) || (m.height & (m.height - 1)))
      FAIL(PBRTB200_EINVAL, "mipmap level 0 must have power-of-two sides (mipmap.rs:162-167)");
    if (m.n_levels < 1 || m.n_levels > 32 || m.wrap > PBRTB200_WRAP_CLAMP) FAIL(PBRTB200_EINVAL, "bad mipmap header");
    uint64_t need = 0;
    for (uint32_t l = 0; l < m.n_levels; ++l)
      need += (uint64_t)std::max(m.width >> l, 1u) * (uint64_t)std::max(m.height >> l, 1u);
    if (!s->texels || m.texel_offset + need > s->n_texels) FAIL(PBRTB200_EINVAL, "mipmap texels out of range");
  }
  for (uint32_t i = 0; i < s->n_meshes; ++i) {
    if (s->n_materials && s->meshes[i].material >= s->n_materials) FAIL(PBRTB200_EINVAL, "mesh material out of range");
    if (s->meshes[i].area_light >= (int32_t)s->n_lights) FAIL(PBRTB200_EINVAL, "mesh area_light out of range");
    if (s->meshes[i].area_light >= 0 && s->lights[s->meshes[i].area_light].kind != PBRTB200_LIGHT_AREA)
      FAIL(PBRTB200_EINVAL, "mesh refers to a non-area light");
  }
  for (uint32_t i = 0; i < s->n_tris; ++i) {
    if (s->tris[i].mesh >= s->n_meshes) FAIL(PBRTB200_EINVAL, "triangle mesh index out of range");
    if ((s->tri_uv || s->tri_n || s->tri_s) && s->tris[i].attr >= s->n_attr)
      FAIL(PBRTB200_EINVAL, "triangle attr index out of range");
  }
  for (uint32_t i = 0; i < s->n_spheres; ++i)
    if (s->n_materials && s->spheres[i].material >= s->n_materials) FAIL(PBRTB200_EINVAL, "sphere material out of range");
  if (s->leaf_prim)
    for (uint32_t i = 0; i < s->n_prims; ++i) {
      const uint32_t p = s->leaf_prim[i];
      if ((p & PB_LEAF_BIT) ? ((p & ~PB_LEAF_BIT) >= s->n_spheres) : (p >= s->n_tris))
        FAIL(PBRTB200_EINVAL, "leaf_prim entry out of range");
    }

  // ---- upload ----
  if (upload(ctx, ctx->d_nodes, pairs.data(), pairs.size())) return PBRTB200_ENODEV;
  if (upload(ctx, ctx->d_tris, s->tris, s->n_tris)) return PBRTB200_ENODEV;
  const bool need_leaf_prim = s->n_spheres > 0 || (s->leaf_prim != nullptr);
  if (need_leaf_prim && upload(ctx, ctx->d_leaf_prim, s->leaf_prim, s->n_prims)) return PBRTB200_ENODEV;
  if (big_leaf && upload(ctx, ctx->d_leaf_count, leaf_count.data(), leaf_count.size())) return PBRTB200_ENODEV;
  const bool need_leaf_boxes = multi || need_leaf_prim;
  if (need_leaf_boxes && upload(ctx, ctx->d_leaf_boxes, pn.leaf_boxes.data(), pn.leaf_boxes.size())) return PBRTB200_ENODEV;
  if (upload(ctx, ctx->d_spheres, s->spheres, s->n_spheres)) return PBRTB200_ENODEV;
  if (upload(ctx, ctx->d_sphere_o2w, s->sphere_o2w, 12ull * s->n_spheres)) return PBRTB200_ENODEV;
  if (upload(ctx, ctx->d_meshes, s->meshes, s->n_meshes)) return PBRTB200_ENODEV;
  if (s->tri_uv && upload(ctx, ctx->d_tri_uv, s->tri_uv, 6ull * s->n_attr)) return PBRTB200_ENODEV;
  if (s->tri_n && upload(ctx, ctx->d_tri_n, s->tri_n, 9ull * s->n_attr)) return PBRTB200_ENODEV;
  if (s->tri_s && upload(ctx, ctx->d_tri_s, s->tri_s, 9ull * s->n_attr)) return PBRTB200_ENODEV;
  if (upload(ctx, ctx->d_materials, s->materials, s->n_materials)) return PBRTB200_ENODEV;
  if (upload(ctx, ctx->d_textures, s->textures, s->n_textures)) return PBRTB200_ENODEV;
  if (s->n_mipmaps) {
    if (upload(ctx, ctx->d_mipmaps, s->mipmaps, s->n_mipmaps)) return PBRTB200_ENODEV;
    if (upload(ctx, ctx->d_texels, s->texels, 4ull * s->n_texels)) return PBRTB200_ENODEV;
  }
  std::vector<uint8_t> mat_flags(std::max<uint32_t>(1u, s->n_materials), 0);
  for (uint32_t i = 0; i < s->n_materials; ++i) {
    const pbrtb200_material& m = s->materials[i];
    auto varying = [&](int id) { return s->textures[id].kind != PBRTB200_TEX_CONSTANT; };
    bool v = varying(m.kd);
    if (m.kind == PBRTB200_MAT_MATTE) v = v || varying(m.sigma);
    if (m.kind == PBRTB200_MAT_PLASTIC) v = v || varying(m.ks) || varying(m.roughness);
    if (m.bump != 0) v = true;  // material::bump reads dudx .. dvdy (material/mod.rs:30-32, 45-47)
    mat_flags[i] = v ? 1 : 0;
  }
  if (upload(ctx, ctx->d_mat_flags, mat_flags.data(), mat_flags.size())) return PBRTB200_ENODEV;
  CK(cudaStreamSynchronize(ctx->stream));  // mat_flags is a local

  DScene& sc = ctx->sc;
  std::memset(&sc, 0, sizeof sc);
  sc.nodes = ctx->d_nodes.as<float4>();
  sc.tris = ctx->d_tris.as<float4>();
  sc.leaf_prim = need_leaf_prim ? ctx->d_leaf_prim.as<uint32_t>() : nullptr;
  sc.leaf_count = big_leaf ? ctx->d_leaf_count.as<uint16_t>() : nullptr;
  sc.leaf_boxes = need_leaf_boxes ? ctx->d_leaf_boxes.as<float4>() : nullptr;
  sc.spheres = ctx->d_spheres.as<pbrtb200_sphere80>();
  sc.sphere_o2w = ctx->d_sphere_o2w.as<float>();
  sc.meshes = ctx->d_meshes.as<pbrtb200_mesh>();
  sc.tri_uv = s->tri_uv ? ctx->d_tri_uv.as<float>() : nullptr;
  sc.tri_n = s->tri_n ? ctx->d_tri_n.as<float>() : nullptr;
  sc.tri_s = s->tri_s ? ctx->d_tri_s.as<float>() : nullptr;
  sc.materials = ctx->d_materials.as<pbrtb200_material>();
  sc.mat_flags = ctx->d_mat_flags.as<uint8_t>();
  sc.textures = ctx->d_textures.as<pbrtb200_texture>();
  sc.mipmaps = s->n_mipmaps ? ctx->d_mipmaps.as<pbrtb200_mipmap>() : nullptr;
  sc.texels = s->n_mipmaps ? ctx->d_texels.as<float4>() : nullptr;
  // PBRTB200_FORCE_EXT=1 runs every scene through the general evaluator (measurement knob: what the
  // out-of-line texture path costs on a scene that does not need it)
  static const bool force_ext = [] {
    const char* v = std::getenv("PBRTB200_FORCE_EXT");
    return v && *v && std::atoi(v) != 0;
  }();
  ctx->shade_ext = force_ext || scene_needs_ext(s);
  sc.n_prims = s->n_prims;
  sc.n_lights = s->n_lights;
  sc.root_ref = pn.root_ref;
  for (int i = 0; i < 3; ++i) {
    sc.root_bmin[i] = pn.root_bmin[i];
    sc.root_bmax[i] = pn.root_bmax[i];
    sc.babs[i] = pn.babs[i];
  }
  sc.boxes_finite = pn.boxes_finite ? 1u : 0u;
  sc.boxes_ordered = pn.boxes_ordered ? 1u : 0u;
  ctx->has_spheres = need_leaf_prim;
  ctx->multi_leaf = multi;

  // ---- lights; area-light triangle records are derived on the device (same arithmetic as a hit
  // on that triangle), their CDF is accumulated here in f32, sequentially. ----
  ctx->h_lights.assign(s->lights, s->lights + s->n_lights);
  sc.light_slots = 0;
  sc.area_sample_pairs = 0;
  for (auto& l : ctx->h_lights) {
    if (l.kind == PBRTB200_LIGHT_AREA) {
      if (l.num_samples < 1) FAIL(PBRTB200_EINVAL, "area light needs num_samples >= 1");
      if (l.n_tris == 0 || (uint64_t)l.first_tri + l.n_tris > s->n_area_prims)
        FAIL(PBRTB200_EINVAL, "area light triangle range invalid");
      sc.light_slots += (uint32_t)l.num_samples;
      sc.area_sample_pairs += (uint32_t)l.num_samples;
    } else if (l.kind == PBRTB200_LIGHT_POINT || l.kind == PBRTB200_LIGHT_SPOT) {
      sc.light_slots += 1;
    } else {
      FAIL(PBRTB200_EINVAL, "unsupported light kind");
    }
  }
  if (s->n_area_prims) {
    for (uint32_t i = 0; i < s->n_area_prims; ++i) {
      const uint32_t p = s->area_prims[i];
      if (p >= s->n_prims || (s->leaf_prim && (s->leaf_prim[p] & PB_LEAF_BIT)))
        FAIL(PBRTB200_EINVAL, "area_prims entry is not a triangle of the primitive list");
    }
    DevBuf d_idx;
    CK(d_idx.ensure(sizeof(uint32_t) * s->n_area_prims));
    CK(cudaMemcpyAsync(d_idx.p, s->area_prims, sizeof(uint32_t) * s->n_area_prims,
                       cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx->d_area_tris.ensure(sizeof(DAreaTri) * s->n_area_prims));
    k_area_tri_setup<<<(s->n_area_prims + 127) / 128, 128, 0, ctx->stream>>>(
        sc, d_idx.as<uint32_t>(), s->n_area_prims, ctx->d_area_tris.as<DAreaTri>());
    CK(cudaGetLastError());
    std::vector<DAreaTri> at(s->n_area_prims);
    CK(cudaMemcpyAsync(at.data(), ctx->d_area_tris.p, sizeof(DAreaTri) * at.size(),
                       cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (auto& l : ctx->h_lights) {
      if (l.kind != PBRTB200_LIGHT_AREA) continue;
      float total = 0.f;
      for (uint32_t k = 0; k < l.n_tris; ++k) total += at[l.first_tri + k].area;
      l.total_area = total;
      float acc = 0.f, lo = 0.f;
      for (uint32_t k = 0; k < l.n_tris; ++k) {
        acc += at[l.first_tri + k].area;
        float hi = acc / total;
        if (k + 1 == l.n_tris) hi = 1.0f;
        at[l.first_tri + k].cdf_lo = lo;
        at[l.first_tri + k].cdf_hi = hi;
        lo = hi;
      }
    }
    CK(cudaMemcpyAsync(ctx->d_area_tris.p, at.data(), sizeof(DAreaTri) * at.size(),
                       cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  if (upload(ctx, ctx->d_lights, ctx->h_lights.data(), ctx->h_lights.size())) return PBRTB200_ENODEV;
  sc.lights = ctx->d_lights.as<pbrtb200_light>();
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->has_scene = true;
  return PBRTB200_OK;
}

static int run_trace_buffer(pbrtb200_ctx* ctx, bool any, const pbrtb200_ray32* d_rays, uint64_t n,
                            pbrtb200_hit16* d_hits, uint8_t* d_occ) {
  CK(cudaMemsetAsync(ctx->d_ctrl.p, 0, sizeof(CtrlBlock), ctx->stream));
  TraceArgs a{};
  a.rays = d_rays;
  a.hits = d_hits;
  a.occluded = d_occ;
  a.n = n;
  a.counter = &ctrl(ctx)->counter;
  a.flags = &ctrl(ctx)->flags;
  DCamera cam{};
  return any ? launch_trace_t<true, 0>(ctx, cam, a) : launch_trace_t<false, 0>(ctx, cam, a);
}

static int finish_flags(pbrtb200_ctx* ctx, CtrlBlock* h) {
  CK(cudaMemcpyAsync(ctx->h_ctrl, ctx->d_ctrl.p, sizeof(CtrlBlock), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *h = *ctx->h_ctrl;
  if (h->flags & 1u) FAIL(PBRTB200_ESTACK, "BVH traversal stack overflow (depth > 64)");
  return 0;
}

int pbrtb200_trace_closest(pbrtb200_ctx* ctx, const pbrtb200_ray32* rays, uint64_t n,
                           pbrtb200_hit16* hits, int is_device, pbrtb200_stats* stats) {
  if (!ctx) return PBRTB200_EINVAL;
  if (!ctx->has_scene) FAIL(PBRTB200_EINVAL, "no scene uploaded");
  if (n && (!rays || !hits)) FAIL(PBRTB200_EINVAL, "rays/hits is NULL");
  CK(cudaSetDevice(ctx->device));
  if (stats) std::memset(stats, 0, sizeof *stats);
  if (n == 0) return PBRTB200_OK;
  const pbrtb200_ray32* d_rays = rays;
  pbrtb200_hit16* d_hits = hits;
  if (!is_device) {
    CK(ctx->d_rays_in.ensure(n * sizeof(pbrtb200_ray32)));
    CK(ctx->d_hits.ensure(n * sizeof(pbrtb200_hit16)));
    CK(cudaMemcpyAsync(ctx->d_rays_in.p, rays, n * sizeof(pbrtb200_ray32), cudaMemcpyHostToDevice, ctx->stream));
    d_rays = ctx->d_rays_in.as<pbrtb200_ray32>();
    d_hits = ctx->d_hits.as<pbrtb200_hit16>();
  }
  StageTimer tm{ctx, stats != nullptr};
  size_t e0 = tm.mark();
  if (int rc = run_trace_buffer(ctx, false, d_rays, n, d_hits, nullptr)) return rc;
  size_t e1 = tm.mark();
  tm.span(e0, e1, 1);
  if (!is_device)
    CK(cudaMemcpyAsync(hits, d_hits, n * sizeof(pbrtb200_hit16), cudaMemcpyDeviceToHost, ctx->stream));
  CtrlBlock h;
  if (int rc = finish_flags(ctx, &h)) return rc;
  if (stats) {
    float ms[6];
    tm.collect(ms);
    stats->camera_rays = n;
    stats->ms_trace = ms[1];
    stats->ms_total = ms[1];
    stats->kernel_launches = 1;
  }
  return PBRTB200_OK;
}

int pbrtb200_trace_any(pbrtb200_ctx* ctx, const pbrtb200_ray32* rays, uint64_t n, uint8_t* occluded,
                       int is_device, pbrtb200_stats* stats) {
  if (!ctx) return PBRTB200_EINVAL;
  if (!ctx->has_scene) FAIL(PBRTB200_EINVAL, "no scene uploaded");
  if (n && (!rays || !occluded)) FAIL(PBRTB200_EINVAL, "rays/occluded is NULL");
  CK(cudaSetDevice(ctx->device));
  if (stats) std::memset(stats, 0, sizeof *stats);
  if (n == 0) return PBRTB200_OK;
  const pbrtb200_ray32* d_rays = rays;
  uint8_t* d_occ = occluded;
  if (!is_device) {
    CK(ctx->d_rays_in.ensure(n * sizeof(pbrtb200_ray32)));
    CK(ctx->d_occ.ensure(n));
    CK(cudaMemcpyAsync(ctx->d_rays_in.p, rays, n * sizeof(pbrtb200_ray32), cudaMemcpyHostToDevice, ctx->stream));
    d_rays = ctx->d_rays_in.as<pbrtb200_ray32>();
    d_occ = ctx->d_occ.as<uint8_t>();
  }
  StageTimer tm{ctx, stats != nullptr};
  size_t e0 = tm.mark();
  if (int rc = run_trace_buffer(ctx, true, d_rays, n, nullptr, d_occ)) return rc;
  size_t e1 = tm.mark();
  tm.span(e0, e1, 3);
  if (!is_device) CK(cudaMemcpyAsync(occluded, d_occ, n, cudaMemcpyDeviceToHost, ctx->stream));
  CtrlBlock h;
  if (int rc = finish_flags(ctx, &h)) return rc;
  if (stats) {
    float ms[6];
    tm.collect(ms);
    stats->shadow_rays = n;
    stats->ms_shadow = ms[3];
    stats->ms_total = ms[3];
    stats->kernel_launches = 1;
  }
  return PBRTB200_OK;
}

// list order -> raster order ([(y - y0) * w + (x - x0)] * spp + i)
__global__ void k_scatter_raster(const DPixel* __restrict__ pixels, uint64_t n_samples, int spp,
                                 int x0, int y0, int w, const float2* __restrict__ img,
                                 const float2* __restrict__ lens, const float* __restrict__ time,
                                 const pbrtb200_hit16* __restrict__ hits, const DCamera cam,
                                 pbrtb200_hit16* __restrict__ out_hits, float* __restrict__ out_samples,
                                 pbrtb200_ray32* __restrict__ out_rays) {
  const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_samples) return;
  const uint64_t p = s / (uint32_t)spp;
  const uint32_t i = (uint32_t)(s - p * (uint32_t)spp);
  const DPixel px = pixels[p];
  const uint64_t o = ((uint64_t)(px_y(px) - y0) * (uint64_t)w + (uint64_t)(px_x(px) - x0)) * (uint32_t)spp + i;
  if (out_hits) out_hits[o] = hits[s];
  const float2 im = img[s];
  const float2 ln = lens ? lens[s] : make_float2(0.f, 0.f);
  if (out_samples) {
    float* q = out_samples + 5 * o;
    q[0] = im.x;
    q[1] = im.y;
    q[2] = ln.x;
    q[3] = ln.y;
    q[4] = time ? time[s] : 0.f;
  }
  if (out_rays) {
    f3 ro, rd;
    camera_ray(cam, im.x, im.y, ln.x, ln.y, &ro, &rd, nullptr);
    pbrtb200_ray32 r;
    r.o[0] = ro.x; r.o[1] = ro.y; r.o[2] = ro.z; r.mint = 0.f;
    r.d[0] = rd.x; r.d[1] = rd.y; r.d[2] = rd.z; r.maxt = PB_F32_MAX;
    out_rays[o] = r;
  }
}

// `film` may be NULL (no edge flags / light floats needed: primary_hits).
static int run_raygen(pbrtb200_ctx* ctx, const DSampler& ds, bool full, uint64_t pix0, uint64_t npix,
                      uint64_t sample0, const pbrtb200_film* film) {
  RaygenArgs ra{};
  ra.pixels = ctx->d_pixels.as<DPixel>() + pix0;
  ra.n_pixels = npix;
  ra.img = ctx->d_img.as<float2>() + sample0;
  ra.lens = full ? ctx->d_lens.as<float2>() + sample0 : nullptr;
  ra.time = full ? ctx->d_time.as<float>() + sample0 : nullptr;
  ra.light_pairs = ctx->sc.area_sample_pairs;
  if (film) {
    ra.lightu = ra.light_pairs ? ctx->d_lightu.as<float2>() + sample0 * ra.light_pairs : nullptr;
    ra.edge = ctx->d_edge.as<uint32_t>() + pix0;
    ra.fx_start = film->x_pixel_start;
    ra.fy_start = film->y_pixel_start;
    ra.fx_count = film->x_pixel_count;
    ra.fy_count = film->y_pixel_count;
    ra.xw = film->filter_xw;
    ra.yw = film->filter_yw;
  }
  if (full) {
    k_raygen_full<<<(unsigned)((npix + 127) / 128), 128, 0, ctx->stream>>>(ds, ra);
  } else {
    const uint64_t nthreads = npix * (uint64_t)((ds.spp + 7) / 8);
    k_raygen_groups<<<(unsigned)((nthreads + 127) / 128), 128, 0, ctx->stream>>>(ds, ra);
  }
  CK(cudaGetLastError());
  return 0;
}

// ---- HaltonSampler (halton.cuh): the padded per-sample layout of the current pixel list ----------
static HaltonArgs halton_args(pbrtb200_ctx* ctx, const pbrtb200_sampler* smp) {
  HaltonArgs a{};
  a.tasks = ctx->d_halton_tasks.as<DHaltonTask>();
  a.n_tasks = ctx->halton_n_tasks;
  a.n_candidates = ctx->halton_candidates;
  a.pix_index = ctx->d_pix_index.as<int32_t>();
  a.sx0 = smp->x_start;
  a.sy0 = smp->y_start;
  a.sw = smp->x_end - smp->x_start;
  a.counts = ctx->d_hcounts.as<uint32_t>();
  a.fill = ctx->d_hcounts.as<uint32_t>() + ctx->n_list_pixels;
  a.offsets = ctx->d_hoffsets.as<uint32_t>();
  a.idx = ctx->d_hidx.as<uint32_t>();
  return a;
}
// Pass 0: accepted candidates per list pixel -> *cap (the largest count, >= 1) and *total.
// Synchronises (the caller sizes its buffers from cap).
static int halton_count(pbrtb200_ctx* ctx, const pbrtb200_sampler* smp, uint32_t* cap, uint64_t* total) {
  if (ctx->halton_valid) {  // same pixel list as the last frame: the binning is still there
    *cap = ctx->halton_cap;
    *total = ctx->halton_total;
    return 0;
  }
  const uint64_t npix = ctx->n_list_pixels;
  CK(ctx->d_hcounts.ensure(2 * npix * sizeof(uint32_t)));  // counts, fill
  CK(cudaMemsetAsync(ctx->d_hcounts.p, 0, 2 * npix * sizeof(uint32_t), ctx->stream));
  const unsigned long long nc = ctx->halton_candidates;
  if (nc >= (1ull << 31) * 256ull) FAIL(PBRTB200_EINVAL, "HaltonSampler: too many candidates for one launch");
  if (nc) {
    k_halton_bin<0><<<(unsigned)((nc + 255) / 256), 256, 0, ctx->stream>>>(halton_args(ctx, smp));
    CK(cudaGetLastError());
  }
  // exclusive scan of the counts on the host: once per pixel list, like the list itself
  std::vector<uint32_t> counts(npix);
  CK(cudaMemcpyAsync(counts.data(), ctx->d_hcounts.p, npix * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->h_hoffsets.assign(npix + 1, 0u);
  uint32_t mx = 0;
  uint64_t sum = 0;
  for (uint64_t i = 0; i < npix; ++i) {
    ctx->h_hoffsets[i] = (uint32_t)sum;
    sum += counts[i];
    mx = std::max(mx, counts[i]);
  }
  if (sum >= 0xFFFFFFFFull) FAIL(PBRTB200_EINVAL, "HaltonSampler: more than 2^32 samples; render in tiles");
  ctx->h_hoffsets[npix] = (uint32_t)sum;
  if (upload(ctx, ctx->d_hoffsets, ctx->h_hoffsets.data(), ctx->h_hoffsets.size())) return PBRTB200_ENODEV;
  CK(cudaStreamSynchronize(ctx->stream));
  *cap = std::max(1u, mx);
  *total = sum;
  ctx->halton_cap = *cap;
  ctx->halton_total = sum;
  return 0;
}
// Passes 1 + 2: scatter the candidate indices, sort per pixel, evaluate the camera samples into
// d_img / d_lens / d_time / d_lightu (halton_total samples each, sized by the caller).
static int halton_fill(pbrtb200_ctx* ctx, const pbrtb200_sampler* smp, bool full, const pbrtb200_film* film) {
  const bool with_film = film != nullptr;
  const uint64_t npix = ctx->n_list_pixels;
  CK(ctx->d_hidx.ensure(std::max<uint64_t>(1, ctx->halton_total) * sizeof(uint32_t)));
  const unsigned long long nc = ctx->halton_candidates;
  if (nc && !ctx->halton_valid) {
    k_halton_bin<1><<<(unsigned)((nc + 255) / 256), 256, 0, ctx->stream>>>(halton_args(ctx, smp));
    CK(cudaGetLastError());
  }
  ctx->halton_valid = true;  // (k_halton_samples sorts each pixel's slots in place: idempotent)
  HaltonSampleArgs sa{};
  sa.tasks = ctx->d_halton_tasks.as<DHaltonTask>();
  sa.pixels = ctx->d_pixels.as<DPixel>();
  sa.n_pixels = npix;
  sa.offsets = ctx->d_hoffsets.as<uint32_t>();
  sa.idx = ctx->d_hidx.as<uint32_t>();
  sa.img = ctx->d_img.as<float2>();
  sa.lens = full ? ctx->d_lens.as<float2>() : nullptr;
  sa.time = full ? ctx->d_time.as<float>() : nullptr;
  sa.light_pairs = with_film ? ctx->sc.area_sample_pairs : 0u;
  sa.lightu = sa.light_pairs ? ctx->d_lightu.as<float2>() : nullptr;
  sa.edge = with_film ? ctx->d_edge.as<uint32_t>() : nullptr;
  sa.xw = with_film ? film->filter_xw : 0.f;
  sa.yw = with_film ? film->filter_yw : 0.f;
  sa.sopen = smp->shutter_open;
  sa.sclose = smp->shutter_close;
  k_halton_samples<<<(unsigned)((npix + 127) / 128), 128, 0, ctx->stream>>>(sa);
  CK(cudaGetLastError());
  return 0;
}

int pbrtb200_halton_layout(pbrtb200_ctx* ctx, const pbrtb200_sampler* smp, uint32_t* cap, uint64_t* n_samples) {
  if (!ctx) return PBRTB200_EINVAL;
  if (int rc = check_sampler(ctx, smp)) return rc;
  if (smp->kind != PBRTB200_SAMPLER_HALTON || !cap) FAIL(PBRTB200_EINVAL, "not a HaltonSampler");
  CK(cudaSetDevice(ctx->device));
  if (int rc = build_pixel_list(ctx, smp, nullptr, nullptr)) return rc;
  uint64_t total = 0;
  if (int rc = halton_count(ctx, smp, cap, &total)) return rc;
  if (n_samples) *n_samples = total;
  return PBRTB200_OK;
}

int pbrtb200_primary_hits(pbrtb200_ctx* ctx, const pbrtb200_camera* cam, const pbrtb200_sampler* smp,
                          pbrtb200_hit16* out_hits, float* out_samples, pbrtb200_ray32* out_rays,
                          int is_device, pbrtb200_stats* stats) {
  if (!ctx) return PBRTB200_EINVAL;
  if (!ctx->has_scene) FAIL(PBRTB200_EINVAL, "no scene uploaded");
  if (!cam) FAIL(PBRTB200_EINVAL, "camera is NULL");
  if (int rc = check_sampler(ctx, smp)) return rc;
  CK(cudaSetDevice(ctx->device));
  if (stats) std::memset(stats, 0, sizeof *stats);
  if (int rc = build_pixel_list(ctx, smp, nullptr, nullptr)) return rc;
  DSampler ds;
  fill_sampler(ctx, smp, &ds);
  DCamera dc;
  fill_camera(cam, ds.spp, &dc);
  const bool halton = smp->kind == PBRTB200_SAMPLER_HALTON;
  uint32_t hcap = 0;
  uint64_t hvalid = 0;
  if (halton)
    if (int rc = halton_count(ctx, smp, &hcap, &hvalid)) return rc;
  // HaltonSampler: compact per-sample buffers (hvalid samples), padded raster layout on output
  const uint64_t npix = ctx->n_list_pixels, ns = halton ? hvalid : npix * (uint64_t)ds.spp;
  const uint64_t n_out = halton ? npix * (uint64_t)hcap : ns;
  const bool full = out_samples != nullptr || cam->lens_radius > 0.0f || smp->kind == PBRTB200_SAMPLER_LD;
  CK(ctx->d_img.ensure(std::max<uint64_t>(1, ns) * sizeof(float2)));
  if (full) {
    CK(ctx->d_lens.ensure(std::max<uint64_t>(1, ns) * sizeof(float2)));
    CK(ctx->d_time.ensure(std::max<uint64_t>(1, ns) * sizeof(float)));
  }
  CK(ctx->d_hits.ensure(std::max<uint64_t>(1, ns) * sizeof(pbrtb200_hit16)));
  StageTimer tm{ctx, stats != nullptr};
  size_t e0 = tm.mark();
  if (halton) {
    if (int rc = halton_fill(ctx, smp, full, nullptr)) return rc;
  } else if (int rc = run_raygen(ctx, ds, full, 0, npix, 0, nullptr)) {
    return rc;
  }
  size_t e1 = tm.mark();
  CK(cudaMemsetAsync(ctx->d_ctrl.p, 0, sizeof(CtrlBlock), ctx->stream));
  TraceArgs a{};
  a.img = ctx->d_img.as<float2>();
  a.lens = (full && cam->lens_radius > 0.0f) ? ctx->d_lens.as<float2>() : nullptr;
  a.hits = ctx->d_hits.as<pbrtb200_hit16>();
  a.n = ns;
  a.counter = &ctrl(ctx)->counter;
  a.flags = &ctrl(ctx)->flags;
  if (int rc = launch_trace_t<false, 1>(ctx, dc, a)) return rc;
  size_t e2 = tm.mark();
  // outputs in raster order
  pbrtb200_hit16* d_oh = nullptr;
  float* d_os = nullptr;
  pbrtb200_ray32* d_or = nullptr;
  if (is_device) {
    d_oh = out_hits;
    d_os = out_samples;
    d_or = out_rays;
  } else {
    if (out_hits) {
      CK(ctx->d_out_a.ensure(n_out * sizeof(pbrtb200_hit16)));
      d_oh = ctx->d_out_a.as<pbrtb200_hit16>();
    }
    if (out_samples) {
      CK(ctx->d_out_b.ensure(n_out * 5 * sizeof(float)));
      d_os = ctx->d_out_b.as<float>();
    }
    if (out_rays) {
      CK(ctx->d_out_c.ensure(n_out * sizeof(pbrtb200_ray32)));
      d_or = ctx->d_out_c.as<pbrtb200_ray32>();
    }
  }
  if (halton && (d_oh || d_os || d_or)) {
    k_scatter_halton<<<(unsigned)((n_out + 255) / 256), 256, 0, ctx->stream>>>(
        ctx->d_pixels.as<DPixel>(), npix, hcap, ctx->d_hoffsets.as<uint32_t>(), smp->x_start, smp->y_start,
        smp->x_end - smp->x_start, ctx->d_img.as<float2>(), full ? ctx->d_lens.as<float2>() : nullptr,
        full ? ctx->d_time.as<float>() : nullptr, ctx->d_hits.as<pbrtb200_hit16>(), dc, d_oh, d_os, d_or);
    CK(cudaGetLastError());
  } else if (d_oh || d_os || d_or) {
    k_scatter_raster<<<(unsigned)((ns + 255) / 256), 256, 0, ctx->stream>>>(
        ctx->d_pixels.as<DPixel>(), ns, ds.spp, smp->x_start, smp->y_start, smp->x_end - smp->x_start,
        ctx->d_img.as<float2>(), full ? ctx->d_lens.as<float2>() : nullptr,
        full ? ctx->d_time.as<float>() : nullptr, ctx->d_hits.as<pbrtb200_hit16>(), dc, d_oh, d_os, d_or);
    CK(cudaGetLastError());
  }
  size_t e3 = tm.mark();
  tm.span(e0, e1, 0);
  tm.span(e1, e2, 1);
  tm.span(e2, e3, 4);
  tm.span(e0, e3, 5);
  if (!is_device) {
    if (out_hits) CK(cudaMemcpyAsync(out_hits, d_oh, n_out * sizeof(pbrtb200_hit16), cudaMemcpyDeviceToHost, ctx->stream));
    if (out_samples) CK(cudaMemcpyAsync(out_samples, d_os, n_out * 5 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (out_rays) CK(cudaMemcpyAsync(out_rays, d_or, n_out * sizeof(pbrtb200_ray32), cudaMemcpyDeviceToHost, ctx->stream));
  }
  CtrlBlock h;
  if (int rc = finish_flags(ctx, &h)) return rc;
  if (stats) {
    float ms[6];
    tm.collect(ms);
    stats->camera_rays = ns;
    stats->ms_raygen = ms[0];
    stats->ms_trace = ms[1];
    stats->ms_film = ms[4];
    stats->ms_total = ms[5];
    stats->kernel_launches = 2 + ((d_oh || d_os || d_or) ? 1 : 0);
  }
  return PBRTB200_OK;
}

int pbrtb200_render(pbrtb200_ctx* ctx, const pbrtb200_camera* cam, const pbrtb200_sampler* smp,
                    const pbrtb200_film* film, const pbrtb200_integrator* integ,
                    const pbrtb200_tileset* tiles, float* out_xyzw, int out_is_device,
                    pbrtb200_stats* stats) {
  if (!ctx) return PBRTB200_EINVAL;
  if (!ctx->has_scene) FAIL(PBRTB200_EINVAL, "no scene uploaded");
  if (!cam || !film || !integ || !out_xyzw) FAIL(PBRTB200_EINVAL, "NULL argument");
  if (int rc = check_sampler(ctx, smp)) return rc;
  if (integ->kind != 0) FAIL(PBRTB200_EINVAL, "unsupported integrator kind");
  if (film->x_pixel_count < 1 || film->y_pixel_count < 1) FAIL(PBRTB200_EINVAL, "empty film");
  if (!(film->filter_xw > 0.f) || !(film->filter_yw > 0.f)) FAIL(PBRTB200_EINVAL, "filter width must be > 0");
  if (ctx->sc.n_lights && !ctx->d_materials.p) FAIL(PBRTB200_EINVAL, "scene has no materials");
  CK(cudaSetDevice(ctx->device));
  if (stats) std::memset(stats, 0, sizeof *stats);
  if (tiles && tiles->n_rects > 0 && !tiles->rects) FAIL(PBRTB200_EINVAL, "tiles->rects is NULL");
  if (tiles && tiles->n_rects == 0) {
    // An EMPTY tile set owns no pixel (a rank whose band is empty): nothing is rendered.  The film is
    // zeroed unless the caller keeps the other owners' pixels — never treated as "the whole film".
    if (!(tiles->flags & PBRTB200_TILES_KEEP_OTHERS)) {
      const size_t bytes = (size_t)film->x_pixel_count * (size_t)film->y_pixel_count * sizeof(float4);
      if (out_is_device) {
        CK(cudaMemsetAsync(out_xyzw, 0, bytes, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
      } else {
        std::memset(out_xyzw, 0, bytes);
      }
    }
    return PBRTB200_OK;
  }
  // PBRTB200_DEBUG_TIMING=1: host wall time of the phases of this call on stderr (each phase ends with a
  // stream synchronisation, so the frame itself gets slower: a diagnostic, e.g. for first-frame latency)
  static const bool dbg_on = [] {
    const char* v = std::getenv("PBRTB200_DEBUG_TIMING");
    return v && *v == '1';
  }();
  auto dbg_t0 = std::chrono::steady_clock::now();
  auto dbg = [&](const char* what) {
    if (!dbg_on) return;
    cudaStreamSynchronize(ctx->stream);
    const auto t1 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[pbrtb200 timing] %-14s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - dbg_t0).count());
    dbg_t0 = t1;
  };
  if (int rc = build_pixel_list(ctx, smp, film, tiles)) return rc;
  dbg("pixel list");

  DSampler ds;
  fill_sampler(ctx, smp, &ds);
  DCamera dc;
  fill_camera(cam, ds.spp, &dc);
  StageTimer tm{ctx, stats != nullptr};
  uint32_t launches = 0;
  size_t eA = tm.mark();
  // HaltonSampler: a variable number of samples per pixel, addressed through per-pixel offsets
  const bool halton = smp->kind == PBRTB200_SAMPLER_HALTON;
  uint32_t hcap = 0;
  uint64_t hvalid = 0;
  if (halton) {
    const bool cached = ctx->halton_valid;
    if (int rc = halton_count(ctx, smp, &hcap, &hvalid)) return rc;
    tm.span(eA, tm.mark(), 0);
    launches += cached ? 0 : 1;  // k_halton_bin<0>
  }
  const int lay = halton ? (int)hcap : ds.spp;  // upper bound of the samples of one list pixel
  const uint64_t npix = ctx->n_list_pixels, ns = halton ? std::max<uint64_t>(1, hvalid) : npix * (uint64_t)ds.spp;
  // first sample of list pixel p (HaltonSampler: host copy of the offsets; else p * spp)
  auto first_sample = [&](uint64_t p) -> uint64_t {
    return halton ? (uint64_t)ctx->h_hoffsets[p] : p * (uint64_t)ds.spp;
  };
  const bool full = cam->lens_radius > 0.0f || smp->kind == PBRTB200_SAMPLER_LD;
  const uint32_t slots = std::max(1u, ctx->sc.light_slots);
  const bool lit = ctx->sc.n_lights > 0;
  const bool multi_slot = lit && slots > 1;  // several radiance terms per sample: folded per chunk (k_fold)
  const uint32_t pairs = ctx->sc.area_sample_pairs;
  if (ctx->sc.n_lights > PB_MAX_FOLD_LIGHTS) FAIL(PBRTB200_EINVAL, "more than 64 lights");
  if (npix >= 0x7FFFFFFFull) FAIL(PBRTB200_EINVAL, "more than 2^31 sampler pixels");

  // ---- wavefront sizing -------------------------------------------------------------------------
  // Chunk-local buffers (hit records, shadow queue, radiance terms of multi-slot scenes) are bounded
  // by kChunkBudgetBytes; shadow-queue slot words index terms with 30 bits.
  const uint64_t chunk_bytes_per_sample = sizeof(pbrtb200_hit16) + (lit ? (uint64_t)slots * (sizeof(pbrtb200_ray32) + sizeof(uint32_t)) : 0) +
                                          (multi_slot ? (uint64_t)slots * sizeof(float4) : 0);
  uint64_t chunk_cap = chunk_samples();
  while (chunk_cap > (1ull << 10) &&
         (chunk_cap * chunk_bytes_per_sample > kChunkBudgetBytes || chunk_cap * slots > PB_SQ_INDEX))
    chunk_cap >>= 1;
  uint64_t chunk_pix = std::max<uint64_t>(1, chunk_cap / (uint64_t)lay);
  chunk_pix = std::min(chunk_pix, npix);
  if (halton && ns <= chunk_cap) chunk_pix = npix;  // the whole frame fits one chunk (lay is only an upper bound per pixel)

  // Per-sample buffers that outlive a chunk (image positions, radiance records, lens / time / light
  // floats) are addressed per list pixel.  A whole-film frame whose buffers would exceed
  // kFrameBudgetBytes keeps them in a RING over list pixels instead: chunks run in list order, after
  // every chunk the film rows that are complete are filtered (k_film per band), and a pixel's slot is
  // reused once no unfiltered film row can reach it.  Memory is then O(chunk), not O(frame).
  const uint64_t frame_bytes_per_sample = sizeof(float2) + (lit ? sizeof(float4) : 0) + (full ? sizeof(float2) + sizeof(float) : 0) +
                                          (uint64_t)pairs * sizeof(float2);
  const bool whole_film = tiles == nullptr;
  const int halo_rows = (int)std::ceil(film->filter_yw + 0.5f) + 2;
  uint64_t ring_px = npix;  // no ring
  if (whole_film && !halton && ns * frame_bytes_per_sample > frame_budget_bytes()) {
    const uint64_t lag_px = (uint64_t)(2 * halo_rows + 8) * (uint64_t)(smp->x_end - smp->x_start);
    uint64_t want = chunk_pix + lag_px, r = 1;
    while (r < want) r <<= 1;
    if (r < npix) ring_px = r;
  }
  const bool ring = ring_px < npix;
  const uint32_t pixel_mask = ring ? (uint32_t)(ring_px - 1) : 0xFFFFFFFFu;
  const uint64_t ring_ns = ring ? ring_px * (uint64_t)lay : ns;
  auto phys = [&](uint64_t p) -> uint64_t { return first_sample(ring ? (p & (ring_px - 1)) : p); };

  uint64_t chunk_ns = chunk_pix * (uint64_t)lay;
  if (halton && chunk_pix == npix) chunk_ns = ns;
  CK(ctx->d_img.ensure(ring_ns * sizeof(float2)));
  if (full) {
    CK(ctx->d_lens.ensure(ring_ns * sizeof(float2)));
    CK(ctx->d_time.ensure(ring_ns * sizeof(float)));
  }
  CK(ctx->d_edge.ensure(npix * sizeof(uint32_t)));
  if (pairs) CK(ctx->d_lightu.ensure(ring_ns * pairs * sizeof(float2)));
  if (lit) CK(ctx->d_rec.ensure(ring_ns * sizeof(float4)));
  if (multi_slot) CK(ctx->d_terms.ensure(chunk_ns * slots * sizeof(float4)));
  CK(ctx->d_hits.ensure(chunk_ns * sizeof(pbrtb200_hit16)));
  if (lit) {
    CK(ctx->d_sq_rays.ensure(chunk_ns * slots * sizeof(pbrtb200_ray32)));
    CK(ctx->d_sq_slots.ensure(chunk_ns * slots * sizeof(uint32_t)));
  }
  const size_t film_px = (size_t)film->x_pixel_count * (size_t)film->y_pixel_count;
  float4* d_film = nullptr;
  // Host film that is page-locked and mapped (cudaHostAlloc, or cudaHostRegister with the Mapped flag
  // as pbrtb200_group_render does for its caller): k_film stores the pixels it owns straight into it
  // over PCIe — the transfer rides under the film kernels (and, for frames of several chunks, under the
  // later chunks) instead of following them as a copy.  Measured on config 3: e2e 8.47 -> 8.35 ms on one
  // GPU, 1.52 -> 1.43 ms on eight.  Pageable buffers, and tile sets that zero the rest of the film, keep
  // the staged copy.  PBRTB200_HOST_FILM_STORES=0 turns it off, =1 limits it to KEEP_OTHERS tile sets.
  static const int host_stores_on = [] {
    const char* v = std::getenv("PBRTB200_HOST_FILM_STORES");
    return v && *v ? std::atoi(v) : 2;
  }();
  bool host_stores = false;
  if (out_is_device) {
    d_film = reinterpret_cast<float4*>(out_xyzw);
  } else {
    if ((host_stores_on >= 1 && tiles && (tiles->flags & PBRTB200_TILES_KEEP_OTHERS)) || (host_stores_on >= 2 && !tiles)) {
      cudaPointerAttributes at{};  // (pageable memory: success, type Unregistered — no error to clear)
      if (cudaPointerGetAttributes(&at, out_xyzw) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
        d_film = reinterpret_cast<float4*>(at.devicePointer);
        host_stores = true;
      } else {
        (void)cudaGetLastError();
      }
    }

    if (!host_stores) {
      CK(ctx->d_film.ensure(film_px * sizeof(float4)));
      d_film = ctx->d_film.as<float4>();
    }
  }

  dbg("buffers");
  CK(cudaMemsetAsync(ctx->d_ctrl.p, 0, sizeof(CtrlBlock), ctx->stream));
  CK(cudaMemsetAsync(ctx->d_edge.p, 0, npix * sizeof(uint32_t), ctx->stream));
  if (tiles && !(tiles->flags & PBRTB200_TILES_KEEP_OTHERS))
    CK(cudaMemsetAsync(d_film, 0, film_px * sizeof(float4), ctx->stream));  // else k_film writes every pixel

  // ---- film stage (launched per band, see below) ---------------------------------------------
  DFilm df;
  df.x_start = film->x_pixel_start;
  df.y_start = film->y_pixel_start;
  df.x_count = film->x_pixel_count;
  df.y_count = film->y_pixel_count;
  df.xw = film->filter_xw;
  df.yw = film->filter_yw;
  df.inv_xw = 1.0f / film->filter_xw;  // filter.rs:12-19
  df.inv_yw = 1.0f / film->filter_yw;
  df.sx0 = smp->x_start;
  df.sx1 = smp->x_end;
  df.sy0 = smp->y_start;
  df.sy1 = smp->y_end;
  df.spp = ds.spp;
  DFold fd{};
  fd.slots = slots;
  fd.n_lights = ctx->sc.n_lights;
  for (uint32_t i = 0; i < ctx->sc.n_lights; ++i) {
    const pbrtb200_light& l = ctx->h_lights[i];
    fd.area[i] = l.kind == PBRTB200_LIGHT_AREA ? 1 : 0;
    fd.ns[i] = (uint16_t)(fd.area[i] ? l.num_samples : 1);
  }
  // k_film runs in BANDS of finished film rows: always with a ring (the samples are about to be
  // overwritten), and for whole-film renders into host memory (each band's device->host copy travels
  // on a second stream while later chunks render).
  static const int band_mode = [] {  // 0: one film launch + one copy; N > 0: N tail pieces
    const char* v = std::getenv("PBRTB200_FILM_BANDS");
    return v && *v ? std::atoi(v) : 2;
  }();
  const bool staged = !out_is_device && !host_stores;  // the film is built in HBM and copied out
  const bool banded = whole_film && !halton && (ring || (band_mode > 0 && staged));
  const bool band_copies = banded && staged;
  struct Band {
    uint32_t y0, y1;
    size_t event;
  };
  std::vector<Band> bands;
  uint32_t film_done = 0;  // film rows [0, film_done) are filtered
  FilmArgs fa{};  // (1 KB of filter table: filled once)
  fa.img = ctx->d_img.as<float2>();
  fa.rec = lit ? ctx->d_rec.as<float4>() : nullptr;
  fa.lights = ctx->sc.lights;
  fa.edge = ctx->d_edge.as<uint32_t>();
  fa.offsets = halton ? ctx->d_hoffsets.as<uint32_t>() : nullptr;
  fa.pix_index = ctx->d_pix_index.as<int32_t>();
  fa.rects = ctx->d_rects.as<int32_t>();
  fa.rect_prefix = ctx->d_rect_prefix.as<uint32_t>();
  fa.n_rects = ctx->n_rects;
  fa.n_pixels = ctx->n_film_pixels;
  fa.pixel_mask = pixel_mask;
  fa.out = d_film;
  std::memcpy(fa.table, film->filter_table, sizeof fa.table);
  auto film_rows = [&](uint32_t y0, uint32_t y1) -> int {
    size_t eF0 = tm.mark();
    if (banded) {  // one rect = the whole film, row-major
      fa.first = y0 * (uint32_t)film->x_pixel_count;
      fa.count = (y1 - y0) * (uint32_t)film->x_pixel_count;
    } else {
      fa.first = 0;
      fa.count = fa.n_pixels;
    }
    k_film<<<(fa.count + 127) / 128, 128, 0, ctx->stream>>>(df, fa);
    CK(cudaGetLastError());
    launches += 1;
    tm.span(eF0, tm.mark(), 4);
    if (banded) {
      if (band_copies) {
        if (bands.size() == ctx->band_events.size()) {
          cudaEvent_t ev;
          CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
          ctx->band_events.push_back(ev);
        }
        CK(cudaEventRecord(ctx->band_events[bands.size()], ctx->stream));
        bands.push_back({y0, y1, bands.size()});
      }
      film_done = y1;
    }
    return 0;
  };
  // k_film's gather extent of film row r, in sampler rows (clamped to the sampler extent)
  auto gather_lo = [&](uint32_t r) -> int {
    const int y = film->y_pixel_start + (int)r;
    return std::max((int)std::ceil((float)y - 0.5f - film->filter_yw) - 1, smp->y_start);
  };
  auto gather_hi = [&](uint32_t r) -> int {
    const int y = film->y_pixel_start + (int)r;
    return std::min((int)std::floor((float)y + 0.5f + film->filter_yw) + 1, smp->y_end - 1);
  };
  // film rows [0, r) are final once `done` list pixels are finished
  auto rows_final = [&](uint64_t done) -> uint32_t {
    uint32_t r = film_done;
    while (r < (uint32_t)film->y_pixel_count) {
      const int qy1 = gather_hi(r);
      if (qy1 >= smp->y_start && (uint64_t)ctx->rows_ready[(size_t)(qy1 - smp->y_start)] > done) break;
      ++r;
    }
    return r;
  };
  // first list pixel an unfiltered film row can still reach (ring: everything before it is free)
  auto live_start = [&]() -> uint64_t {
    if (film_done >= (uint32_t)film->y_pixel_count) return npix;
    const int qy0 = gather_lo(film_done);
    return ctx->row_first[(size_t)std::min(std::max(qy0 - smp->y_start, 0), smp->y_end - smp->y_start)];
  };

  if (halton) {  // candidates land anywhere: every sample of the frame is generated up front
    size_t h0 = tm.mark();
    const bool cached = ctx->halton_valid;
    if (int rc = halton_fill(ctx, smp, full, film)) return rc;
    tm.span(h0, tm.mark(), 0);
    launches += cached ? 1 : 2;  // (k_halton_bin<1>,) k_halton_samples
  }
  for (uint64_t p0 = 0; p0 < npix;) {
    uint64_t cp = std::min(chunk_pix, npix - p0);
    if (ring) {
      cp = std::min(cp, ring_px - (p0 & (ring_px - 1)));  // a chunk is contiguous in the ring
      const uint64_t room = ring_px - (p0 - live_start());
      if (room == 0) FAIL(PBRTB200_ENOMEM, "sample ring too small for this filter width (PBRTB200_CHUNK_LOG2)");
      cp = std::min(cp, room);
    }
    const uint64_t s0 = phys(p0);
    const uint64_t cn = halton ? first_sample(p0 + cp) - first_sample(p0) : cp * (uint64_t)lay;
    const uint64_t p1 = p0 + cp;
    if (cn == 0) {  // (HaltonSampler: a run of pixels without a single sample)
      p0 = p1;
      continue;
    }
    size_t e0 = tm.mark();
    if (!halton)
      if (int rc = run_raygen(ctx, ds, full, p0, cp, s0, film)) return rc;
    size_t e1 = tm.mark();
    dbg("raygen");
    CK(cudaMemsetAsync(&ctrl(ctx)->counter, 0, sizeof(unsigned long long), ctx->stream));
    CK(cudaMemsetAsync(&ctrl(ctx)->sq_count, 0, sizeof(uint32_t), ctx->stream));
    TraceArgs ta{};
    ta.img = ctx->d_img.as<float2>() + s0;
    ta.lens = full && cam->lens_radius > 0.0f ? ctx->d_lens.as<float2>() + s0 : nullptr;
    ta.hits = ctx->d_hits.as<pbrtb200_hit16>();
    ta.n = cn;
    ta.counter = &ctrl(ctx)->counter;
    ta.flags = &ctrl(ctx)->flags;
    if (tiles && !halton) {  // halo pixels that cannot reach an owned pixel are not traced
      ta.pixels = ctx->d_pixels.as<DPixel>() + p0;
      ta.edge = ctx->d_edge.as<uint32_t>() + p0;
      ta.spp = (uint32_t)ds.spp;
    }
    if (int rc = launch_trace_t<false, 1>(ctx, dc, ta)) return rc;
    size_t e2 = tm.mark();
    dbg("closest hit");
    launches += 2;
    tm.span(e0, e1, 0);
    tm.span(e1, e2, 1);
    if (lit) {
      ShadeArgs sa{};
      sa.img = ta.img;
      sa.lens = ta.lens;
      sa.lightu = pairs ? ctx->d_lightu.as<float2>() + s0 * pairs : nullptr;
      sa.hits = ta.hits;
      sa.area_tris = ctx->d_area_tris.as<DAreaTri>();
      // one slot: the terms ARE the radiance records; several: chunk-local terms, folded below
      sa.terms = multi_slot ? ctx->d_terms.as<float4>() : ctx->d_rec.as<float4>() + s0;
      sa.sq_rays = ctx->d_sq_rays.as<pbrtb200_ray32>();
      sa.sq_slots = ctx->d_sq_slots.as<uint32_t>();
      sa.sq_count = &ctrl(ctx)->sq_count;
      sa.hit_total = &ctrl(ctx)->hit_total;
      sa.nan_count = &ctrl(ctx)->nan_count;
      sa.n = cn;
      sa.slots = slots;
      sa.strict_flags = integ->strict_flags;
      if (ctx->shade_ext)
        k_shade<PB_SHADE_EXT_MIN_BLOCKS, true, true><<<(unsigned)((cn + 127) / 128), 128, 0, ctx->stream>>>(ctx->sc, dc, sa);
      else if (ctx->sc.mipmaps)
        k_shade<PB_SHADE_MIN_BLOCKS, true><<<(unsigned)((cn + 127) / 128), 128, 0, ctx->stream>>>(ctx->sc, dc, sa);
      else
        k_shade<PB_SHADE_MIN_BLOCKS, false><<<(unsigned)((cn + 127) / 128), 128, 0, ctx->stream>>>(ctx->sc, dc, sa);
      CK(cudaGetLastError());
      size_t e3 = tm.mark();
      dbg("shade");
      CK(cudaMemsetAsync(&ctrl(ctx)->counter, 0, sizeof(unsigned long long), ctx->stream));
      TraceArgs sh{};
      sh.rays = sa.sq_rays;
      sh.contrib = sa.terms;
      sh.slots = sa.sq_slots;
      sh.nan_count = &ctrl(ctx)->nan_count;
      sh.n_dyn = &ctrl(ctx)->sq_count;
      sh.shadow_total = &ctrl(ctx)->shadow_total;
      sh.counter = &ctrl(ctx)->counter;
      sh.flags = &ctrl(ctx)->flags;
      if (int rc = launch_trace_t<true, 0>(ctx, dc, sh)) return rc;
      size_t e4 = tm.mark();
      dbg("any hit");
      tm.span(e2, e3, 2);
      tm.span(e3, e4, 3);
      launches += 2;
      if (multi_slot) {
        FoldArgs fo{};
        fo.terms = sa.terms;
        fo.lights = ctx->sc.lights;
        fo.rec = ctx->d_rec.as<float4>() + s0;
        fo.n = cn;
        fo.nan_count = &ctrl(ctx)->nan_count;
        k_fold<<<(unsigned)((cn + 255) / 256), 256, 0, ctx->stream>>>(fd, fo);
        CK(cudaGetLastError());
        tm.span(e4, tm.mark(), 2);
        launches += 1;
      }
    }
    if (banded && p1 < npix) {
      const uint32_t r = rows_final(p1);
      if (r > film_done)
        if (int rc = film_rows(film_done, r)) return rc;
    }
    p0 = p1;
  }
  if (!banded) {
    if (int rc = film_rows(0u, (uint32_t)film->y_pixel_count)) return rc;
  } else {
    // the rows left after the last chunk go out in a few pieces so that the copy of one piece
    // overlaps the filtering of the next
    const uint32_t H = (uint32_t)film->y_pixel_count;
    const uint32_t pieces = (uint32_t)std::max(1, band_copies ? band_mode : 1);
    const uint32_t step = std::max(16u, (H - film_done + pieces - 1u) / pieces);
    while (film_done < H)
      if (int rc = film_rows(film_done, std::min(H, film_done + step))) return rc;
  }
  tm.span(eA, tm.mark(), 5);
  dbg("film");
  if (staged) {
    if (band_copies) {
      // every band: wait for its film launch on the copy stream, then move its rows to the host
      for (const Band& b : bands) {
        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->band_events[b.event], 0));
        const size_t off = (size_t)b.y0 * (size_t)film->x_pixel_count;
        CK(cudaMemcpyAsync(out_xyzw + 4 * off, d_film + off,
                           (size_t)(b.y1 - b.y0) * (size_t)film->x_pixel_count * sizeof(float4),
                           cudaMemcpyDeviceToHost, ctx->copy_stream));
      }
      CK(cudaStreamSynchronize(ctx->copy_stream));
    } else if (tiles && (tiles->flags & PBRTB200_TILES_KEEP_OTHERS)) {
      // only the rects this call owns travel to the host film (a group's devices each send their own
      // rows over their own PCIe link); the rest of the caller's buffer is left alone
      for (uint32_t q = 0; q < tiles->n_rects; ++q) {
        const int32_t* rc4 = tiles->rects + 4 * q;
        const size_t x0 = (size_t)(rc4[0] - film->x_pixel_start), y0r = (size_t)(rc4[1] - film->y_pixel_start);
        const size_t wr = (size_t)(rc4[2] - rc4[0]), hr = (size_t)(rc4[3] - rc4[1]);
        const size_t pitch = (size_t)film->x_pixel_count * sizeof(float4), off = y0r * (size_t)film->x_pixel_count + x0;
        CK(cudaMemcpy2DAsync(out_xyzw + 4 * off, pitch, d_film + off, pitch, wr * sizeof(float4), hr,
                             cudaMemcpyDeviceToHost, ctx->stream));
      }
    } else {
      CK(cudaMemcpyAsync(out_xyzw, d_film, film_px * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    }
  }
  CtrlBlock h;
  if (int rc = finish_flags(ctx, &h)) return rc;
  dbg("copy + finish");
  if (stats) {
    float ms[6];
    tm.collect(ms);
    stats->camera_rays = halton ? hvalid : ns;
    stats->camera_hits = h.hit_total;
    stats->shadow_rays = h.shadow_total;
    stats->ms_raygen = ms[0];
    stats->ms_trace = ms[1];
    stats->ms_shade = ms[2];
    stats->ms_shadow = ms[3];
    stats->ms_film = ms[4];
    stats->ms_total = ms[5];
    stats->kernel_launches = launches;
    stats->nan_samples = h.nan_count;
  }
  if (h.nan_count) FAIL(PBRTB200_ENAN, "Invalid radiance value!");
  return PBRTB200_OK;
}


int pbrtb200_cost_profile(pbrtb200_ctx* ctx, const pbrtb200_camera* cam, const pbrtb200_film* film, int stride,
                          float* row_cost) {
  if (!ctx) return PBRTB200_EINVAL;
  if (!ctx->has_scene) FAIL(PBRTB200_EINVAL, "no scene uploaded");
  if (!cam || !film || !row_cost || stride < 1) FAIL(PBRTB200_EINVAL, "bad argument");
  if (film->x_pixel_count < 1 || film->y_pixel_count < 1) FAIL(PBRTB200_EINVAL, "empty film");
  CK(cudaSetDevice(ctx->device));
  const int h = film->y_pixel_count;
  CK(ctx->d_out_b.ensure((size_t)h * sizeof(float)));
  CK(cudaMemsetAsync(ctx->d_out_b.p, 0, (size_t)h * sizeof(float), ctx->stream));
  DCamera dc;
  fill_camera(cam, 1, &dc);
  dc.lens_radius = 0.0f;  // (probe rays are pinhole rays)
  TraceLaunchCfg cfg{ctx->stream, ctx->sm_count, ctx->has_spheres, ctx->multi_leaf};
  // constant part of a camera sample (raygen, shading, film) in units of traversal steps: on
  // config 3 those stages take about a third of the time of its ~55 steps per sample
  const float base_cost = 20.0f;
  CK(pb_launch_cost_probe(cfg, ctx->sc, dc, film->x_pixel_start, film->y_pixel_start, film->x_pixel_count, h, stride,
                          base_cost, ctx->d_area_tris.p, ctx->d_out_b.as<float>()));
  CK(cudaMemcpyAsync(row_cost, ctx->d_out_b.p, (size_t)h * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  for (int y = 0; y < h; ++y)  // a probe row stands for the `stride` rows it starts
    if (y % stride) row_cost[y] = row_cost[y - y % stride];
  return PBRTB200_OK;
}

int pbrtb200_film_develop(pbrtb200_ctx* ctx, const float* film_xyzw, int film_is_device,
                          uint64_t n_pixels, float* out_rgb, uint8_t* out_rgb8, int out_is_device) {
  if (!ctx) return PBRTB200_EINVAL;
  if (n_pixels && !film_xyzw) FAIL(PBRTB200_EINVAL, "film is NULL");
  if (!out_rgb && !out_rgb8) FAIL(PBRTB200_EINVAL, "no output requested");
  CK(cudaSetDevice(ctx->device));
  if (n_pixels == 0) return PBRTB200_OK;
  const float4* d_in = reinterpret_cast<const float4*>(film_xyzw);
  if (!film_is_device) {
    CK(ctx->d_film.ensure(n_pixels * sizeof(float4)));
    CK(cudaMemcpyAsync(ctx->d_film.p, film_xyzw, n_pixels * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    d_in = ctx->d_film.as<float4>();
  }
  float* d_rgb = out_rgb;
  uint8_t* d_rgb8 = out_rgb8;
  if (!out_is_device) {
    if (out_rgb) {
      CK(ctx->d_out_b.ensure(n_pixels * 3 * sizeof(float)));
      d_rgb = ctx->d_out_b.as<float>();
    }
    if (out_rgb8) {
      CK(ctx->d_out_c.ensure(n_pixels * 3));
      d_rgb8 = ctx->d_out_c.as<uint8_t>();
    }
  }
  k_film_develop<<<(unsigned)((n_pixels + 255) / 256), 256, 0, ctx->stream>>>(d_in, n_pixels, d_rgb, d_rgb8);
  CK(cudaGetLastError());
  if (!out_is_device) {
    if (out_rgb) CK(cudaMemcpyAsync(out_rgb, d_rgb, n_pixels * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (out_rgb8) CK(cudaMemcpyAsync(out_rgb8, d_rgb8, n_pixels * 3, cudaMemcpyDeviceToHost, ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return PBRTB200_OK;
}

int pbrtb200_peer_film_create(pbrtb200_ctx* ctx, uint64_t n_pixels, void** dev_ptr, unsigned char handle64[64]) {
  if (!ctx || !dev_ptr || !handle64 || n_pixels == 0) return PBRTB200_EINVAL;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  CK(cudaSetDevice(ctx->device));
  // a dedicated cudaMalloc allocation per call: IPC handles name whole allocations
  void* p = nullptr;
  CK(cudaMalloc(&p, n_pixels * sizeof(float4)));
  ctx->peer_films.push_back(p);
  CK(cudaMemsetAsync(p, 0, n_pixels * sizeof(float4), ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, p));
  std::memcpy(handle64, &h, 64);
  *dev_ptr = p;
  return PBRTB200_OK;
}

int pbrtb200_peer_film_open(pbrtb200_ctx* ctx, const unsigned char handle64[64], void** dev_ptr) {
  if (!ctx || !dev_ptr || !handle64) return PBRTB200_EINVAL;
  CK(cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, 64);
  void* p = nullptr;
  CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *dev_ptr = p;
  return PBRTB200_OK;
}

int pbrtb200_peer_film_close(pbrtb200_ctx* ctx, void* dev_ptr) {
  if (!ctx || !dev_ptr) return PBRTB200_EINVAL;
  CK(cudaSetDevice(ctx->device));
  CK(cudaIpcCloseMemHandle(dev_ptr));
  return PBRTB200_OK;
}
}  // extern "C"
