// Traversal kernels, box-test family 2 (trace_core.cuh child_box).
#define PB_TRACE_BOX 2
#include "trace_launch.inl"
