// Closest-hit / any-hit BVH traversal kernels (sm_100a).
//
// Replaces, per ray: BVHAccelerator::intersect (src/primitive/aggregates/bvh.rs:375-422),
// BBox::intersect (src/bbox.rs:185-209), Triangle::get_intersection_point (src/shape/mesh.rs:41-72),
// Sphere::get_intersection_point (src/shape/sphere.rs:46-107) and GeometricPrimitive::intersect's
// ray.maxt update (src/primitive/geometric.rs:62-73).
//
// Equivalence with the reference's test-at-pop order (proved in DESIGN.md §"Traversal"): the
// reference pops a node and tests its box against the LIVE [mint,maxt]; the test passes iff
// T0 <= F and T0 <= maxt_now with T0 = max(mint, near_xyz), F = min(far_xyz) (both independent of
// maxt).  We evaluate both children's boxes when visiting the parent, push the far child together
// with its T0 only if it passes then, and re-check T0 <= maxt when it is popped.  Because maxt
// only shrinks, the set and ORDER of visited leaves — hence every accepted hit, including the
// "equal t replaces" tie rule (mesh.rs:71, bvh.rs:401-404) — is identical.
#pragma once
#include "scene.cuh"
#include "trace_math.cuh"  // slab_test, slab_test_finite, tri_hit, camera_ray
#include "trace_core.cuh"  // sphere_hit, traverse, trace_ray (one ray through the pair-node BVH)

// Flags in the upper bits of a shadow-queue slot word (chunk-local term index in the low 30 bits).
#define PB_SQ_NAN 0x80000000u    // the guarded term holds a NaN: an unoccluded ray makes the sample's radiance NaN
#define PB_SQ_KEEPW 0x40000000u  // the term's w carries an emitter index: zero xyz only
#define PB_SQ_INDEX 0x3FFFFFFFu

// ---- kernels ---------------------------------------------------------------------------------
// Persistent grid: each warp pulls 32-ray packets from a global counter (warp-aggregated: one
// atomic per warp), so slow packets do not stall a whole block's worth of queued work.

struct TraceArgs {
  const pbrtb200_ray32* __restrict__ rays;  // SRC 0
  const float2* __restrict__ img;           // SRC 1: camera samples (image_x, image_y)
  const float2* __restrict__ lens;          // SRC 1, may be NULL (lens_radius == 0)
  pbrtb200_hit16* __restrict__ hits;        // closest-hit output (may be NULL for ANY)
  uint8_t* __restrict__ occluded;           // ANY output (API hook), may be NULL
  float4* __restrict__ contrib;             // ANY in the render pipeline: zero slot if occluded
  const uint32_t* __restrict__ slots;       // ANY in the render pipeline: term index | flags per ray
  uint32_t* nan_count;                      // ANY in the render pipeline: unoccluded rays guarding a NaN term
  const uint32_t* __restrict__ n_dyn;       // if non-NULL the ray count is read from device memory
  unsigned long long* shadow_total;         // if non-NULL: += ray count (stats)
  uint64_t n;
  unsigned long long* counter;  // work counter, zeroed by the host before launch
  uint32_t* flags;              // bit0: stack overflow happened
  uint32_t batch;               // 32-ray packets a warp claims per atomic (>= 1)
  // SRC 1, tiled renders: a halo pixel whose samples all stay inside their own pixel (edge flag 0,
  // set by raygen with add_sample's own arithmetic) cannot reach a pixel this call owns, so its
  // rays are not traced (hit = MISS).  With the box filter that is practically every halo pixel.
  const DPixel* __restrict__ pixels;   // may be NULL: trace everything
  const uint32_t* __restrict__ edge;
  uint32_t spp;
};

#ifndef PB_TRACE_MIN_BLOCKS
#define PB_TRACE_MIN_BLOCKS(ANY) 12  // 40 registers: occupancy pays (profiles/r02_notes.md)
#endif
template <bool ANY, bool SPH, bool MULTI, int SRC, int MODE, int BOX>
__global__ void __launch_bounds__(PB_TRACE_THREADS, PB_TRACE_MIN_BLOCKS(ANY))
k_trace(const DScene sc, const DCamera cam, const TraceArgs a) {
  // child refs, then (closest hit only: any-hit keeps no T0) the entry distances
  constexpr int SMD = PB_SM_STACK_OF(ANY);
  __shared__ uint32_t sh_stack[(ANY ? 1 : 2) * (SMD > 0 ? SMD : 1) * (SMD > 0 ? PB_TRACE_THREADS : 1)];
  uint32_t* s_ref = sh_stack + (SMD > 0 ? threadIdx.x : 0);
  float* s_t0 = reinterpret_cast<float*>(s_ref + (ANY ? 0 : SMD * PB_TRACE_THREADS));
  const int lane = threadIdx.x & 31;
  const uint64_t n = a.n_dyn ? (uint64_t)(*a.n_dyn) : a.n;
  if (a.shadow_total && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(a.shadow_total, (unsigned long long)n);
  const unsigned long long claim = 32ull * a.batch;
  for (;;) {
    unsigned long long base0 = 0;
    if (lane == 0) base0 = atomicAdd(a.counter, claim);
    base0 = __shfl_sync(0xffffffffu, base0, 0);
    if (base0 >= n) break;  // warp-uniform exit
   for (unsigned long long base = base0; base < base0 + claim && base < n; base += 32ull) {
    const uint64_t idx = base + lane;
    if (idx < n) {
      f3 o, d;
      float mint, maxt;
      if (SRC == 0) {
        const float4* rp = reinterpret_cast<const float4*>(a.rays + idx);
        const float4 r0 = ld_stream(rp), r1 = ld_stream(rp + 1);
        o = mk3(r0.x, r0.y, r0.z);
        mint = r0.w;
        d = mk3(r1.x, r1.y, r1.z);
        maxt = r1.w;
      } else {
        if (a.pixels) {
          const uint64_t pix = idx / a.spp;
          if ((__ldg(&a.pixels[pix].task) & PB_PIXEL_HALO_BIT) && __ldg(&a.edge[pix]) == 0u) {
            st_stream(reinterpret_cast<float4*>(a.hits) + idx, make_float4(__uint_as_float(PBRTB200_MISS), 0.f, 0.f, 0.f));
            continue;
          }
        }
        const float2 im = ld_stream(a.img + idx);
        float2 ln = make_float2(0.f, 0.f);
        if (a.lens) ln = ld_stream(a.lens + idx);
        camera_ray(cam, im.x, im.y, ln.x, ln.y, &o, &d, nullptr);
        mint = 0.0f;         // ray.rs:30-38 Ray::new_with(.., start = 0)
        maxt = PB_F32_MAX;
      }
      TraceResult r = trace_ray<ANY, SPH, MULTI, MODE, BOX>(sc, o, d, mint, maxt, s_ref, s_t0);
      if (r.prim == PB_OVERFLOW) {
        atomicOr(a.flags, 1u);
        r.prim = PBRTB200_MISS;
      }
      if (ANY) {
        const bool occ = r.prim != PBRTB200_MISS;
        if (a.occluded) a.occluded[idx] = occ ? 1 : 0;
        if (a.contrib) {  // render pipeline: the guarded radiance term (shade.cuh PB_SQ_*)
          const uint32_t sl = ld_stream(a.slots + idx);
          float4* term = a.contrib + (sl & PB_SQ_INDEX);
          if (occ) {
            if (sl & PB_SQ_KEEPW) {  // w carries an emitter index
              term->x = 0.f;
              term->y = 0.f;
              term->z = 0.f;
            } else {
              *term = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          } else if (sl & PB_SQ_NAN) {
            atomicAdd(a.nan_count, 1u);  // sampler_renderer.rs:105 intent: counted once, where it becomes final
          }
        }
      } else {
        const uint64_t oi = idx;
        float4 h;
        h.x = __uint_as_float(r.prim);
        h.y = r.t;
        h.z = r.b1;
        h.w = r.b2;
        st_stream(reinterpret_cast<float4*>(a.hits) + oi, h);
      }
    }
   }
  }
}


