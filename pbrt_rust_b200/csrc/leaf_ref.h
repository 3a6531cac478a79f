// Encoding of a BVH child reference (shared by the device kernels and the host-side packing).
#pragma once
#define PB_LEAF_BIT 0x80000000u
// Leaf ref = bit31 | (min(count,16)-1) << 27 | prim_offset (27 bits: up to 134 M primitives).
// A count field of 15 means "16 or more: read leaf_count[prim_offset]".
#define PB_LEAF_CNT_SHIFT 27
#define PB_LEAF_OFF_MASK 0x07FFFFFFu
