// Selects the traversal-kernel family (trace_launch.h).
#include "trace_launch.h"

cudaError_t pb_launch_trace(bool any, int src, int box, const TraceLaunchCfg& cfg, const DScene& sc,
                            const DCamera& cam, const TraceArgs& a) {
  switch (box) {
    case 1: return pb_launch_trace_box1(any, src, cfg, sc, cam, a);
    case 2: return pb_launch_trace_box2(any, src, cfg, sc, cam, a);
    default: return pb_launch_trace_box3(any, src, cfg, sc, cam, a);
  }
}
