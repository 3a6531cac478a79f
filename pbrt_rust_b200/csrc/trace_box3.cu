// Traversal kernels, box-test family 3 (trace_core.cuh child_box).
#define PB_TRACE_BOX 3
#include "trace_launch.inl"
