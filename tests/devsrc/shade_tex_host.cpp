// TEST INFRASTRUCTURE ONLY.  Compiles the device source pbrt_rust_b200/csrc/shade_tex.cuh as host
// code (PB_HOST_CHECK: no CUDA, intrinsics replaced by plain C++) so the CPU test-suite can run the
// very arithmetic the GPU kernel executes against the oracle where no GPU exists.  The product
// never loads this library; image textures (MIPMap lookups) stay device-only and are not covered.
#define PB_HOST_CHECK 1
#include "../../pbrt_rust_b200/csrc/shade_tex.cuh"

#include <cstring>

static_assert(sizeof(DG) == 30 * sizeof(float), "DG is 30 packed floats");

extern "C" {
// dg30 = struct DG: p, nn, u, v, dpdu, dpdv, dndu, dndv, dpdx, dpdy, dudx, dudy, dvdx, dvdy
void devsrc_tex_eval(const pbrtb200_texture* table, int id, const float* dg30, float* out3) {
  DG dg;
  std::memcpy(&dg, dg30, sizeof dg);
  TexEnv env{table, nullptr, nullptr};
  const f3 r = tex_eval_ext<PBRTB200_TEX_MAX_DEPTH>(env, id, dg);
  out3[0] = r.x;
  out3[1] = r.y;
  out3[2] = r.z;
}
void devsrc_tex_map(const pbrtb200_texture* tx, const float* dg30, float* out6) {
  DG dg;
  std::memcpy(&dg, dg30, sizeof dg);
  tex_map_ext(*tx, dg, out6);
}
// out9 = bumped dpdu, dpdv, nn
void devsrc_bump(const pbrtb200_texture* table, int id, const float* dg30, const float* ng3, int flip,
                 float* out9) {
  DG dg;
  std::memcpy(&dg, dg30, sizeof dg);
  TexEnv env{table, nullptr, nullptr};
  const DG b = bump_dg_(env, id, dg, mk3(ng3[0], ng3[1], ng3[2]), flip != 0);
  out9[0] = b.dpdu.x; out9[1] = b.dpdu.y; out9[2] = b.dpdu.z;
  out9[3] = b.dpdv.x; out9[4] = b.dpdv.y; out9[5] = b.dpdv.z;
  out9[6] = b.nn.x;   out9[7] = b.nn.y;   out9[8] = b.nn.z;
}
float devsrc_noise(float x, float y, float z) { return noise_(x, y, z); }
float devsrc_fbm(int turb, const float* p, const float* dpdx, const float* dpdy, float omega, int octaves) {
  return fbm_(turb != 0, mk3(p[0], p[1], p[2]), mk3(dpdx[0], dpdx[1], dpdx[2]), mk3(dpdy[0], dpdy[1], dpdy[2]),
              omega, octaves);
}
}

// ---- HaltonSampler arithmetic (pbrt_rust_b200/csrc/halton_math.cuh) -------------------------------
#include "../../pbrt_rust_b200/csrc/halton_math.cuh"
extern "C" {
double devsrc_radical_inverse(unsigned long long n, unsigned int b) { return radical_inverse_(n, b); }
// win4 = x0, x1, y0, y1 of the task window; returns 1 and the image position when the candidate is kept
int devsrc_halton_image(const int* win4, float delta, unsigned long long i, float* out2) {
  DHaltonTask t{};
  t.x0 = win4[0]; t.x1 = win4[1]; t.y0 = win4[2]; t.y1 = win4[3];
  t.delta = delta;
  return halton_image(t, i, &out2[0], &out2[1]) ? 1 : 0;
}
unsigned int devsrc_halton_prime(int k) { return pb_halton_primes[k]; }
}

// ---- traversal arithmetic (csrc/trace_math.cuh) and dmath.cuh helpers ------------------------------
#include "../../pbrt_rust_b200/csrc/trace_math.cuh"
extern "C" {
// box6 = bmin, bmax; ray8 = o, mint, d, maxt.  finite != 0: the min/max variant for finite 1/d.
int devsrc_slab(const float* box6, const float* ray8, int finite, float* t0_out) {
  const f3 o = mk3(ray8[0], ray8[1], ray8[2]);
  const f3 inv = mk3(1.f / ray8[4], 1.f / ray8[5], 1.f / ray8[6]);
  return (finite ? slab_test_finite(box6[0], box6[1], box6[2], box6[3], box6[4], box6[5], o, inv, ray8[3], ray8[7], t0_out)
                 : slab_test(box6[0], box6[1], box6[2], box6[3], box6[4], box6[5], o, inv, ray8[3], ray8[7], t0_out))
             ? 1 : 0;
}
int devsrc_tri_hit(const float* p9, const float* ray8, float* tbb) {
  return tri_hit(mk3(p9[0], p9[1], p9[2]), mk3(p9[3], p9[4], p9[5]), mk3(p9[6], p9[7], p9[8]), mk3(ray8[0], ray8[1], ray8[2]),
                 mk3(ray8[4], ray8[5], ray8[6]), ray8[3], ray8[7], &tbb[0], &tbb[1], &tbb[2]) ? 1 : 0;
}
int devsrc_quadratic(float a, float b, float c, float* t01) { return quadratic_(a, b, c, &t01[0], &t01[1]) ? 1 : 0; }
int devsrc_solve2x2(const float* a4, const float* b2, float* x2) {
  return solve2x2_(a4[0], a4[1], a4[2], a4[3], b2[0], b2[1], &x2[0], &x2[1]) ? 1 : 0;
}
void devsrc_coordinate_system(const float* v3, float* out6) {
  f3 a, b;
  coordinate_system_(mk3(v3[0], v3[1], v3[2]), &a, &b);
  out6[0] = a.x; out6[1] = a.y; out6[2] = a.z; out6[3] = b.x; out6[4] = b.y; out6[5] = b.z;
}
void devsrc_chacha12_block(const uint32_t* key8, unsigned long long blk, uint32_t* out16) { chacha12_block(key8, blk, out16); }
}
