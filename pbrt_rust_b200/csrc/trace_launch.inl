// Body of one traversal-kernel family; included by trace_box<N>.cu with PB_TRACE_BOX = N.
#include <cstdio>

#include "trace_launch.h"

// SIMT loop shapes (trace_core.cuh traverse): measured best per kernel kind
#ifndef PB_MODE_ANY
#define PB_MODE_ANY 2
#endif
#ifndef PB_MODE_CLOSEST
#define PB_MODE_CLOSEST 1
#endif
namespace {
template <class K>
int grid_for(K kfn, int sm_count) {
  static int per_sm = 0;  // one static per kernel instantiation
  if (per_sm == 0) {
    int v = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kfn, PB_TRACE_THREADS, 0) != cudaSuccess || v < 1) {
      (void)cudaGetLastError();
      v = 4;
    }
    per_sm = v;
  }
  return sm_count * per_sm;
}
template <bool ANY, int SRC>
cudaError_t launch(const TraceLaunchCfg& cfg, const DScene& sc, const DCamera& cam, const TraceArgs& a) {
  constexpr int MODE = ANY ? PB_MODE_ANY : PB_MODE_CLOSEST;
#define PB_LAUNCH(SPH, MULTI)                                                      \
  {                                                                                \
    auto kfn = k_trace<ANY, SPH, MULTI, SRC, MODE, PB_TRACE_BOX>;                  \
    kfn<<<grid_for(kfn, cfg.sm_count), PB_TRACE_THREADS, 0, cfg.stream>>>(sc, cam, a); \
  }
  if (cfg.spheres) {
    if (cfg.multi) PB_LAUNCH(true, true) else PB_LAUNCH(true, false)
  } else {
    if (cfg.multi) PB_LAUNCH(false, true) else PB_LAUNCH(false, false)
  }
#undef PB_LAUNCH
  return cudaGetLastError();
}
}  // namespace

#define PB_CAT2(a, b) a##b
#define PB_CAT(a, b) PB_CAT2(a, b)
cudaError_t PB_CAT(pb_launch_trace_box, PB_TRACE_BOX)(bool any, int src, const TraceLaunchCfg& cfg, const DScene& sc,
                                                      const DCamera& cam, const TraceArgs& a) {
  if (any) return launch<true, 0>(cfg, sc, cam, a);
  return src ? launch<false, 1>(cfg, sc, cam, a) : launch<false, 0>(cfg, sc, cam, a);
}
