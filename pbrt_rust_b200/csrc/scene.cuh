// Device-side scene layout (HBM) for the B200 back end.  See DESIGN.md §"Data layout in HBM".
#pragma once
#include "../../include/pbrtb200.h"
#include "dmath.cuh"

#include "leaf_ref.h"  // PB_LEAF_BIT, PB_LEAF_CNT_SHIFT, PB_LEAF_OFF_MASK
// Traversal-stack entries kept in shared memory, per kernel kind; deeper entries (all of them for a
// depth of 0) live in local memory, i.e. in L1.  Shared memory and L1 share one 256 KB array per SM:
// the closest-hit kernel keeps a short shared stack (24 entries cost it ~150 KB of L1 and 8 % of its
// speed), the any-hit kernel none at all (measured, profiles/r02_notes.md).
#ifndef PB_SM_STACK_CLOSEST
#define PB_SM_STACK_CLOSEST 8
#endif
#ifndef PB_SM_STACK_ANY
#define PB_SM_STACK_ANY 0
#endif
#define PB_SM_STACK_OF(ANY) ((ANY) ? PB_SM_STACK_ANY : PB_SM_STACK_CLOSEST)
#define PB_SM_STACK_MAX (PB_SM_STACK_CLOSEST > PB_SM_STACK_ANY ? PB_SM_STACK_CLOSEST : PB_SM_STACK_ANY)
#ifndef PB_TRACE_THREADS
#define PB_TRACE_THREADS 128
#endif

// BVH "pair node", 64 B = 4 x float4, built at upload from the reference's linear PackedBVHNode
// array (pbrtb200_node32).  An inner node stores BOTH children's boxes so one 64-byte fetch
// (4 x LDG.128) replaces the two dependent 32-byte fetches of a test-at-pop traversal:
//   q0 = (c0.min.x, c0.min.y, c0.min.z, c0.max.x)
//   q1 = (c0.max.y, c0.max.z, c1.min.x, c1.min.y)
//   q2 = (c1.min.z, c1.max.x, c1.max.y, c1.max.z)
//   q3 = bits(ref0, ref1, 1 << axis, axis)
// c0 is the reference's first child (node i+1), c1 its second child (second_child_offset).
// ref: bit31 = 0 -> index of another pair node; bit31 = 1 -> leaf, low 31 bits = prim_offset into
// the ordered primitive list plus an inline primitive count (see PB_LEAF_CNT_SHIFT).
struct DScene {
  const float4* __restrict__ nodes;
  const float4* __restrict__ tris;            // 3 x float4 per triangle (pbrtb200_tri48)
  const uint32_t* __restrict__ leaf_prim;     // NULL when the scene has no spheres (prim i == tri i)
  const uint16_t* __restrict__ leaf_count;    // NULL when every leaf holds exactly one primitive
  const float4* __restrict__ leaf_boxes;      // 2 x float4 per primitive offset: the box of the leaf that
                                              // starts there, as the reference stores it (min.xyz, max.xyz);
                                              // NULL when every leaf is one triangle whose box equals the
                                              // bounds of its vertices (checked at upload)
  const pbrtb200_sphere80* __restrict__ spheres;
  const float* __restrict__ sphere_o2w;
  const pbrtb200_mesh* __restrict__ meshes;
  const float* __restrict__ tri_uv;
  const float* __restrict__ tri_n;
  const float* __restrict__ tri_s;
  const pbrtb200_material* __restrict__ materials;
  const uint8_t* __restrict__ mat_flags;  // bit0: some texture of the material is not Constant
  const pbrtb200_texture* __restrict__ textures;
  const pbrtb200_light* __restrict__ lights;
  const pbrtb200_mipmap* __restrict__ mipmaps;  // image textures: headers ...
  const float4* __restrict__ texels;            // ... and the texel pool (rgb, 0)
  uint32_t root_ref;
  uint32_t n_prims;
  uint32_t n_lights;
  uint32_t light_slots;        // sum over lights of (area ? num_samples : 1)
  uint32_t area_sample_pairs;  // sum over area lights of num_samples (RNG pairs per camera sample)
  float root_bmin[3], root_bmax[3];
  float babs[3];           // max |coordinate| over every node box, per axis (BOX 3 error bound)
  uint32_t boxes_finite;   // every node box coordinate is finite (else: compare-and-swap tests only)
  uint32_t boxes_ordered;  // every node box has min <= max on every axis (octant-specialised tests)
};

// Internal record of one emissive triangle (built on device at upload from scene.area_prims).
struct DAreaTri {
  float p1[3], p2[3], p3[3], nn[3];
  float area, cdf_lo, cdf_hi, pad;
};

struct DCamera {
  float r2c[16];
  float c2w[16];
  float dx[3], dy[3];
  float sopen, sclose, lens_radius, focal_distance;
  float diff_scale;  // 1 / sqrt(samples_per_pixel)  (sampler_renderer.rs:96)
};

// One sampler pixel to evaluate: position, owning task and index within the task's sub-window
// (which fixes its offset in the task's RNG word stream).
struct DPixel {
  int32_t xy;     // x | y << 16 (signed 16-bit each; sample extents can start at -radius)
  uint32_t k;     // raster index of the pixel inside its task window
  uint32_t task;  // task index -> key; bit 31: halo pixel (needed for the filter footprint of an
                  // owned tile, not itself inside a tile this call renders)
};
#define PB_PIXEL_HALO_BIT 0x80000000u
#define PB_PIXEL_TASK_MASK 0x7FFFFFFFu
PB_DEV int px_x(DPixel p) { return (int)(short)(p.xy & 0xFFFF); }
PB_DEV int px_y(DPixel p) { return (int)(short)((p.xy >> 16) & 0xFFFF); }

struct DSampler {
  int kind;  // 0 stratified, 1 LD
  int xs, ys, jitter;
  int spp;
  uint32_t words_per_pixel;
  uint32_t cam_words;  // words of the camera-sample block (light-sample floats follow)
  float sopen, sclose;
  const uint32_t* __restrict__ task_keys;  // 8 words per task
};
