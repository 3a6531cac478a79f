// Cost probe (sm_100a): one camera ray through the centre of every `stride`-th film pixel, plus one
// shadow ray per light from its hit point, each counted in traversal steps.  The per-row sums are a
// cheap proxy of where a frame's time goes; pbrtb200_group_render cuts the film into row bands of
// equal cost with it, so that a ONE-SHOT multi-GPU render is balanced before any frame has been
// timed (the reference balances by work stealing over its task queue, sampler_renderer.rs:168-173).
// Not on the parity path: it never touches a pixel value.
#include "trace_launch.h"

namespace {
struct ProbeArgs {
  int x0, y0, w, h, stride;   // film pixel extent and probe stride
  float* row_cost;            // h floats, zeroed by the caller
  float base_cost;            // constant cost of a camera sample outside the traversal (steps)
  const DAreaTri* area_tris;
};

template <bool SPH, bool MULTI>
__global__ void __launch_bounds__(PB_TRACE_THREADS)
k_cost_probe(const DScene sc, const DCamera cam, const ProbeArgs a) {
  constexpr int SMD = PB_SM_STACK_OF(false);
  __shared__ uint32_t sh_stack[2 * (SMD > 0 ? SMD : 1) * (SMD > 0 ? PB_TRACE_THREADS : 1)];
  uint32_t* s_ref = sh_stack + (SMD > 0 ? threadIdx.x : 0);
  float* s_t0 = reinterpret_cast<float*>(s_ref + SMD * PB_TRACE_THREADS);
  const int pw = (a.w + a.stride - 1) / a.stride, ph = (a.h + a.stride - 1) / a.stride;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pw * ph) return;
  const int px = (i % pw) * a.stride, py = (i / pw) * a.stride;
  f3 o, d;
  camera_ray(cam, (float)(a.x0 + px) + 0.5f, (float)(a.y0 + py) + 0.5f, 0.5f, 0.5f, &o, &d, nullptr);
  const TraceResult r = trace_ray<false, SPH, MULTI, 0, 1, true>(sc, o, d, 0.0f, PB_F32_MAX, s_ref, s_t0);
  float cost = a.base_cost + (float)r.steps;
  if (r.prim != PBRTB200_MISS && r.prim != PB_OVERFLOW) {
    const f3 p = o + d * r.t;
    for (uint32_t li = 0; li < sc.n_lights; ++li) {
      const pbrtb200_light lt = sc.lights[li];
      f3 target = mk3(lt.pos[0], lt.pos[1], lt.pos[2]);
      float ns = 1.f;
      if (lt.kind == PBRTB200_LIGHT_AREA) {  // centroid of the light's first emissive triangle
        const DAreaTri at = a.area_tris[lt.first_tri];
        target = mk3((at.p1[0] + at.p2[0] + at.p3[0]) / 3.f, (at.p1[1] + at.p2[1] + at.p3[1]) / 3.f,
                     (at.p1[2] + at.p2[2] + at.p3[2]) / 3.f);
        ns = (float)lt.num_samples;
      }
      const f3 w = target - p;
      const float dist = len3(w);
      if (!(dist > 0.f)) continue;
      // (the any-hit loop shapes keep no T0: the closest-hit stack layout of this kernel is a superset)
      const TraceResult s = trace_ray<true, SPH, MULTI, 0, 1, true>(sc, p, w * (1.f / dist), r.t * 5e-4f, dist * 0.999f, s_ref, s_t0);
      cost += ns * (a.base_cost * 0.25f + (float)s.steps);
    }
  }
  // every probe stands for stride rows x stride columns
  atomicAdd(&a.row_cost[py], cost);
}
}  // namespace

cudaError_t pb_launch_cost_probe(const TraceLaunchCfg& cfg, const DScene& sc, const DCamera& cam, int x0, int y0,
                                 int w, int h, int stride, float base_cost, const void* area_tris, float* d_row_cost) {
  ProbeArgs a{x0, y0, w, h, stride, d_row_cost, base_cost, reinterpret_cast<const DAreaTri*>(area_tris)};
  const int pw = (w + stride - 1) / stride, ph = (h + stride - 1) / stride;
  const unsigned grid = (unsigned)((pw * ph + PB_TRACE_THREADS - 1) / PB_TRACE_THREADS);
  if (cfg.spheres) {
    if (cfg.multi) k_cost_probe<true, true><<<grid, PB_TRACE_THREADS, 0, cfg.stream>>>(sc, cam, a);
    else k_cost_probe<true, false><<<grid, PB_TRACE_THREADS, 0, cfg.stream>>>(sc, cam, a);
  } else {
    if (cfg.multi) k_cost_probe<false, true><<<grid, PB_TRACE_THREADS, 0, cfg.stream>>>(sc, cam, a);
    else k_cost_probe<false, false><<<grid, PB_TRACE_THREADS, 0, cfg.stream>>>(sc, cam, a);
  }
  return cudaGetLastError();
}
