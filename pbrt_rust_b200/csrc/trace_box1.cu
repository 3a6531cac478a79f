// Traversal kernels, box-test family 1 (trace_core.cuh child_box).
#define PB_TRACE_BOX 1
#include "trace_launch.inl"
