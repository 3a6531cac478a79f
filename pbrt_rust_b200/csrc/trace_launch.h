// Host-side entry to the traversal kernels (csrc/trace.cuh).  The kernels live in their own
// translation units (trace_box*.cu, one per box-test family) so that they compile in parallel with
// the rest of the library; api.cu only sees this declaration.
#pragma once
#include <cuda_runtime.h>

#include "trace.cuh"

struct TraceLaunchCfg {
  cudaStream_t stream;
  int sm_count;
  bool spheres;  // the scene has quadrics / an indirection list (leaf_prim)
  bool multi;    // some leaf holds more than one primitive (or leaf boxes must be read from memory)
};

// any: any-hit (VisibilityTester) instead of closest hit.  src: 0 = rays from a buffer, 1 = camera
// samples (closest hit only).  box: 1, 2, 3 = box-test family (trace_core.cuh child_box).
// Loop shapes are fixed to the measured best: closest hit while-while (MODE 1), any-hit if-if
// without child ordering (MODE 2); profiles/r01_notes.md.
cudaError_t pb_launch_trace(bool any, int src, int box, const TraceLaunchCfg& cfg, const DScene& sc,
                            const DCamera& cam, const TraceArgs& a);
// one family each (trace_box1.cu, trace_box2.cu, trace_box3.cu)
cudaError_t pb_launch_trace_box1(bool any, int src, const TraceLaunchCfg& cfg, const DScene& sc, const DCamera& cam, const TraceArgs& a);
cudaError_t pb_launch_trace_box2(bool any, int src, const TraceLaunchCfg& cfg, const DScene& sc, const DCamera& cam, const TraceArgs& a);
cudaError_t pb_launch_trace_box3(bool any, int src, const TraceLaunchCfg& cfg, const DScene& sc, const DCamera& cam, const TraceArgs& a);

// Cost probe (trace_probe.cu): per film row, the summed traversal cost (node steps + primitive tests,
// camera ray + one shadow ray per light) of one probe ray per stride x stride pixels.
cudaError_t pb_launch_cost_probe(const TraceLaunchCfg& cfg, const DScene& sc, const DCamera& cam, int x0, int y0,
                                 int w, int h, int stride, float base_cost, const void* area_tris, float* d_row_cost);
