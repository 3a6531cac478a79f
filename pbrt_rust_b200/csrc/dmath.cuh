// Device math for the B200 back end.  Every function keeps the floating-point operation ORDER of
// the pbrt_rust function it replaces (cited per function, paths relative to the reference root);
// the translation unit is compiled with -fmad=false -prec-div=true -prec-sqrt=true so that nvcc
// neither contracts a*b+c into an FMA nor substitutes approximate div/sqrt — Rust does neither,
// and primary-hit primitive ids are required to match bit-exactly.
#pragma once
#include <stdint.h>
#ifdef PB_HOST_CHECK
// Host compilation of the device arithmetic for the CPU test-suite (tests/devsrc/): the few CUDA
// intrinsics used below get plain C++ equivalents.  Never part of the product build.
#include <algorithm>
#include <cmath>
#include <cstring>
using std::isnan;
using std::max;
using std::min;
#define PB_DEV inline
struct float4 {
  float x, y, z, w;
};
struct float2 {
  float x, y;
};
inline int __float2int_rz(float x) {  // cvt.rzi.s32.f32: saturating, NaN -> 0
  if (std::isnan(x)) return 0;
  if (x >= 2147483648.0f) return INT32_MAX;
  if (x <= -2147483648.0f) return INT32_MIN;
  return (int)x;
}
inline uint32_t __float_as_uint(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  return u;
}
inline float __uint_as_float(uint32_t u) {
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
inline float __int_as_float(int i) { return __uint_as_float((uint32_t)i); }
// Kernels under the host check: one emulated thread per block (blockDim = 1), which is what the
// block-level bookkeeping of k_shade degenerates to with a single lane; the harness sets blockIdx.
struct PbDim3 {
  unsigned x, y, z;
};
static PbDim3 blockIdx{0, 0, 0}, threadIdx{0, 0, 0};
static const PbDim3 blockDim{1, 1, 1};
#define __global__
#define __shared__ static
#define __launch_bounds__(...)
inline unsigned __ballot_sync(unsigned, bool p) { return p ? 1u : 0u; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline void __syncthreads() {}
inline unsigned __activemask() { return 1u; }
inline bool __all_sync(unsigned, bool p) { return p; }  // a single lane always agrees with itself
inline int __ffs(int v) { return __builtin_ffs(v); }
template <class T>
inline T __shfl_sync(unsigned, T v, int) {
  return v;
}
template <class T, class U>
inline T atomicOr(T* p, U v) {
  const T old = *p;
  *p = (T)(old | (T)v);
  return old;
}
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline float2 make_float2(float x, float y) { return float2{x, y}; }
template <class T, class U>
inline T atomicAdd(T* p, U v) {  // one host thread: a plain read-modify-write
  const T old = *p;
  *p = (T)(old + (T)v);
  return old;
}
template <class T>
inline T __ldg(const T* p) {  // ld.global.nc
  return *p;
}
inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, int c) {
  c &= 31;
  return c ? (hi << c) | (lo >> (32 - c)) : hi;
}
#else
#include <cuda_runtime.h>
#define PB_DEV __device__ __forceinline__
#endif
#ifdef PB_HOST_CHECK
#define PB_PREFETCH_L1(p) ((void)(p))
#define PB_PREFETCH_L2(p) ((void)(p))
#else
#define PB_PREFETCH_L1(p) asm volatile("prefetch.global.L1 [%0];" ::"l"(p))
#define PB_PREFETCH_L2(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
#endif
// Streaming accesses: the per-sample wavefront buffers (image samples, hit records, shadow rays, radiance
// terms: 0.5 - 1 GB each per config-3 frame) are written once and read once, a kernel later.
// -DPB_STREAM_HINTS=1 marks them evict-first (ld / st .cs).  Measured (profiles/r02_notes.md §12): the L2
// hit rates and DRAM bytes of the traversal kernels do not move — their DRAM reads ARE the streamed
// samples / rays, the scene already stays in L2 — and k_shade gets slower (config 3 1.338 -> 1.361 ms,
// config 5 63 -> 75 ms).  Off.
#ifndef PB_STREAM_HINTS
#define PB_STREAM_HINTS 0
#endif
#ifdef PB_HOST_CHECK
template <class T> PB_DEV T ld_stream(const T* p) { return *p; }
template <class T> PB_DEV void st_stream(T* p, const T& v) { *p = v; }
#elif PB_STREAM_HINTS
template <class T> PB_DEV T ld_stream(const T* p) { return __ldcs(p); }
template <class T> PB_DEV void st_stream(T* p, const T& v) { __stcs(p, v); }
#else
template <class T> PB_DEV T ld_stream(const T* p) { return __ldg(p); }
template <class T> PB_DEV void st_stream(T* p, const T& v) { *p = v; }
#endif
#define PB_F32_MAX 3.402823466e+38f
#define PB_PI 3.14159265358979323846f

struct f3 {
  float x, y, z;
};
PB_DEV f3 mk3(float x, float y, float z) {
  f3 r;
  r.x = x;
  r.y = y;
  r.z = z;
  return r;
}
// geometry/vector.rs:45-130 (component-wise; f*v == v*f)
PB_DEV f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
PB_DEV f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
PB_DEV f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
PB_DEV f3 operator*(f3 a, float f) { return mk3(a.x * f, a.y * f, a.z * f); }
PB_DEV f3 operator*(float f, f3 a) { return mk3(a.x * f, a.y * f, a.z * f); }
// vector.rs:112-126: v / f == (1/f) * v
PB_DEV f3 operator/(f3 a, float f) {
  float recip = 1.0f / f;
  return a * recip;
}
// vector.rs:168-172
PB_DEV float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// vector.rs:174-181
PB_DEV f3 cross3(f3 a, f3 v) {
  return mk3((a.y * v.z) - (a.z * v.y), (a.z * v.x) - (a.x * v.z), (a.x * v.y) - (a.y * v.x));
}
PB_DEV float len2(f3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
PB_DEV float len3(f3 a) { return sqrtf(len2(a)); }
// geometry/normal.rs:180-185
PB_DEV f3 normalize3(f3 a) { return a / len3(a); }
PB_DEV float comp(f3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

// Rust f32::clamp (NaN stays NaN)
PB_DEV float rclampf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }
// Rust `f32 as i32` (saturating, NaN -> 0) — __float2int_rz saturates and maps NaN to 0.
PB_DEV int f2i_sat(float x) { return __float2int_rz(x); }
// utils/mod.rs:16-21
PB_DEV float lerpf_(float a, float b, float t) { return a * (1.0f - t) + b * t; }

// vector.rs:184-195 as written
PB_DEV void coordinate_system_(f3 v1, f3* o1, f3* o2) {
  f3 v2;
  if (fabsf(v1.x) > fabsf(v1.y)) {
    float inv_len = 1.0f / sqrtf(v1.x * v1.x + v1.z * v1.z);
    v2 = mk3(-v1.x * inv_len, 0.f, v1.x * inv_len);
  } else {
    float inv_len = 1.0f / sqrtf(v1.y * v1.y + v1.z * v1.z);
    v2 = mk3(0.f, v1.z * inv_len, -v1.y * inv_len);
  }
  f3 v3 = cross3(v1, v2);
  *o1 = cross3(v3, v1);
  *o2 = v3;
}

// Affine 3x4 (rows 0..2 of a 4x4 whose last row is 0 0 0 1).  transform/transform.rs:207-239.
// For such matrices w == 1 exactly, so the Point path never divides.
PB_DEV f3 xf_pt(const float* m, f3 p) {
  return mk3(m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7],
             m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]);
}
PB_DEV f3 xf_vec(const float* m, f3 v) {
  return mk3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z,
             m[8] * v.x + m[9] * v.y + m[10] * v.z);
}
// Normal: transpose of the inverse (rows of m_inv indexed by column)
PB_DEV f3 xf_nrm(const float* mi, f3 n) {
  return mk3(mi[0] * n.x + mi[4] * n.y + mi[8] * n.z, mi[1] * n.x + mi[5] * n.y + mi[9] * n.z,
             mi[2] * n.x + mi[6] * n.y + mi[10] * n.z);
}
// Full 4x4 Point transform with the homogeneous divide (raster_to_camera is projective).
PB_DEV f3 xf_pt44(const float* m, f3 p) {
  float xt = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3];
  float yt = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7];
  float zt = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11];
  float w = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15];
  if (w != 1.f) return mk3(xt / w, yt / w, zt / w);
  return mk3(xt, yt, zt);
}

// utils/mod.rs:56-92
PB_DEV bool quadratic_(float a, float b, float c, float* t0, float* t1) {
  float descrim = b * b - 4.f * a * c;
  if (descrim < 0.0f) return false;
  if (fabsf(descrim) < 1e-6f) {
    if (a == 0.0f) return false;
    float t = -b / (2.0f * a);
    *t0 = t;
    *t1 = t;
    return true;
  }
  float root_descrim = sqrtf(descrim);
  float q = (b < 0.0f) ? -0.5f * (b - root_descrim) : -0.5f * (b + root_descrim);
  if (a == 0.0f) return false;
  float x0 = q / a;
  float x1 = c / q;
  if (x0 < x1) {
    *t0 = x0;
    *t1 = x1;
  } else {
    *t0 = x1;
    *t1 = x0;
  }
  return true;
}

// utils/mod.rs:94-110
PB_DEV bool solve2x2_(float a00, float a01, float a10, float a11, float b0, float b1, float* x0,
                      float* x1) {
  float det = a00 * a11 - a01 * a10;
  if (fabsf(det) < 1e-10f) return false;
  float inv_det = 1.0f / det;
  float r0 = (a11 * b0 - a01 * b1) * inv_det;
  float r1 = (a00 * b1 - a10 * b0) * inv_det;
  if (isnan(r0) || isnan(r1)) return false;
  *x0 = r0;
  *x1 = r1;
  return true;
}

// ---- ChaCha12 (rand 0.8 StdRng = rand_chacha 0.3 ChaCha12Rng; see DESIGN.md "RNG") ----------
PB_DEV uint32_t rotl32_(uint32_t v, int c) { return __funnelshift_l(v, v, c); }
#define PB_QR(a, b, c, d)  \
  a += b;                  \
  d = rotl32_(d ^ a, 16);  \
  c += d;                  \
  b = rotl32_(b ^ c, 12);  \
  a += b;                  \
  d = rotl32_(d ^ a, 8);   \
  c += d;                  \
  b = rotl32_(b ^ c, 7);
// Output block `blk` of the stream keyed by key[8] (64-bit counter in words 12..13, stream id 0).
PB_DEV void chacha12_block(const uint32_t* __restrict__ key, uint64_t blk, uint32_t out[16]) {
  const uint32_t c0 = 0x61707865u, c1 = 0x3320646eu, c2 = 0x79622d32u, c3 = 0x6b206574u;
  uint32_t k0 = key[0], k1 = key[1], k2 = key[2], k3 = key[3], k4 = key[4], k5 = key[5],
           k6 = key[6], k7 = key[7];
  uint32_t n0 = (uint32_t)blk, n1 = (uint32_t)(blk >> 32);
  uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = c3, x4 = k0, x5 = k1, x6 = k2, x7 = k3, x8 = k4,
           x9 = k5, x10 = k6, x11 = k7, x12 = n0, x13 = n1, x14 = 0, x15 = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    PB_QR(x0, x4, x8, x12)
    PB_QR(x1, x5, x9, x13)
    PB_QR(x2, x6, x10, x14)
    PB_QR(x3, x7, x11, x15)
    PB_QR(x0, x5, x10, x15)
    PB_QR(x1, x6, x11, x12)
    PB_QR(x2, x7, x8, x13)
    PB_QR(x3, x4, x9, x14)
  }
  out[0] = x0 + c0;
  out[1] = x1 + c1;
  out[2] = x2 + c2;
  out[3] = x3 + c3;
  out[4] = x4 + k0;
  out[5] = x5 + k1;
  out[6] = x6 + k2;
  out[7] = x7 + k3;
  out[8] = x8 + k4;
  out[9] = x9 + k5;
  out[10] = x10 + k6;
  out[11] = x11 + k7;
  out[12] = x12 + n0;
  out[13] = x13 + n1;
  out[14] = x14;
  out[15] = x15;
}
// rng.rs:15-17: gen::<f32>() = (u32 >> 8) * 2^-24
PB_DEV float u32_to_unit_float(uint32_t w) { return (float)(w >> 8) * (1.0f / 16777216.0f); }

// Sequential reader over the ChaCha12 word stream (block cache in local memory).
struct WordStream {
  const uint32_t* key;
  uint64_t pos;
  uint64_t cached;
  uint32_t buf[16];
  PB_DEV void init(const uint32_t* k, uint64_t p) {
    key = k;
    pos = p;
    cached = ~0ull;
  }
  PB_DEV uint32_t next_u32() {
    uint64_t blk = pos >> 4;
    if (blk != cached) {
      chacha12_block(key, blk, buf);
      cached = blk;
    }
    uint32_t w = buf[pos & 15];
    pos++;
    return w;
  }
  PB_DEV float random_float() { return u32_to_unit_float(next_u32()); }
  // rng.rs:19-21: gen::<u64>() % usize::MAX
  PB_DEV uint64_t random_uint() {
    uint64_t lo = next_u32();
    uint64_t hi = next_u32();
    uint64_t v = lo | (hi << 32);
    return v % 0xFFFFFFFFFFFFFFFFull;
  }
};

// ---- low-discrepancy helpers and the RNG shuffle (used by k_raygen_full) ------------------------
// sampler/utils.rs:6-20 (as written: the last bit-reversal step shifts by 2)
PB_DEV float van_der_corput_(uint32_t n, uint32_t scramble) {
  n = (n << 16) | (n >> 16);
  n = ((n & 0x00ff00ffu) << 8) | ((n & 0xff00ff00u) >> 8);
  n = ((n & 0x0f0f0f0fu) << 4) | ((n & 0xf0f0f0f0u) >> 4);
  n = ((n & 0x33333333u) << 2) | ((n & 0xCCCCCCCCu) >> 2);
  n = ((n & 0x55555555u) << 2) | ((n & 0xAAAAAAAAu) >> 2);
  n ^= scramble;
  return (float)((double)((n >> 8) & 0xffffffu) / 16777216.0);
}
// sampler/utils.rs:22-35
PB_DEV float sobol2_(uint32_t n, uint32_t scramble) {
  uint32_t s = scramble, v = 1u << 31;
  while (n != 0) {
    if ((n & 1u) == 0) s ^= v;
    v ^= v >> 1;
    n >>= 1;
  }
  return (float)((double)((s >> 8) & 0xFFFFFFu) / 16777216.0);
}

// rng.rs:23-33 over `count` groups of DIMS floats stored as float / float2 in global memory
PB_DEV void shuffle2(WordStream& ws, float2* v, uint32_t count) {
  for (uint32_t i = 0; i < count; ++i) {
    const uint32_t other = i + (uint32_t)(ws.random_uint() % (uint64_t)(count - i));
    const float2 t = v[i];
    v[i] = v[other];
    v[other] = t;
  }
}
PB_DEV void shuffle1(WordStream& ws, float* v, uint32_t count) {
  for (uint32_t i = 0; i < count; ++i) {
    const uint32_t other = i + (uint32_t)(ws.random_uint() % (uint64_t)(count - i));
    const float t = v[i];
    v[i] = v[other];
    v[other] = t;
  }
}
