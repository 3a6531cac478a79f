/*
 * pbrtb200.h — C ABI of the B200 rendering back end for pbrt_rust.
 *
 * This is the drop-in boundary.  The reference-side interface it replaces is
 *     trait Renderer { fn render(&mut self, scene: &Scene); ... }      (src/renderer.rs:8-26)
 * as implemented by SamplerRenderer (src/sampler_renderer.rs:26-54, 147-182): a host crate
 * `GpuRenderer: Renderer` flattens its Scene/Camera/Sampler/Film into the plain structs below and
 * calls pbrtb200_render(); see INTEGRATION.md for the Rust `extern "C"` block and the flatten shim.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; pbrtb200_last_error(ctx) gives the text
 *     (panics of the reference become error codes: NaN radiance — sampler_renderer.rs:105 intent;
 *     singular matrix — transform/matrix4x4.rs:157; traversal-stack overflow).
 *   - the caller owns every input array; the library copies during the call and keeps no host
 *     pointer.  Device memory is owned by the ctx.  Output buffers are caller-allocated.
 *   - a ctx is bound to one CUDA device, is not thread-safe, and every call is synchronous.
 *   - all structs are POD with fixed layout (`#[repr(C)]` on the Rust side).
 *   - there is NO CPU fallback: without a CUDA device every entry point fails with PBRTB200_ENODEV.
 */
#ifndef PBRTB200_H
#define PBRTB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBRTB200_OK 0
#define PBRTB200_EINVAL (-1)    /* bad argument / malformed scene */
#define PBRTB200_ENODEV (-2)    /* no CUDA device / CUDA failure */
#define PBRTB200_ENAN (-3)      /* "Invalid radiance value!" (sampler_renderer.rs:105) */
#define PBRTB200_ESTACK (-4)    /* traversal stack deeper than PBRTB200_STACK_DEPTH */
#define PBRTB200_ESINGULAR (-5) /* "Singular matrix!" (matrix4x4.rs:157) */
#define PBRTB200_ENOMEM (-6)

#define PBRTB200_STACK_DEPTH 64 /* bvh.rs:386 Vec::with_capacity(64) is a hint; here a hard bound */
#define PBRTB200_MISS 0xFFFFFFFFu

typedef struct pbrtb200_ctx pbrtb200_ctx;

/* ---- flattened scene --------------------------------------------------------------------- */

/* One PackedBVHNode (src/primitive/aggregates/bvh.rs:260-272), in the reference's depth-first
 * order: the first child of inner node i is i+1, the second is `offset`.                        */
typedef struct {
  float bmin[3];
  float bmax[3];
  uint32_t offset;  /* Leaf: prim_offset into the ordered primitive list; Inner: second child  */
  uint16_t count;   /* Leaf: num_prims (1..65535); Inner: 0                                     */
  uint8_t axis;     /* Inner: split axis 0..2                                                   */
  uint8_t is_leaf;  /* 1 = Leaf, 0 = Inner                                                      */
} pbrtb200_node32;

/* One Triangle (src/shape/mesh.rs:27-39) with world-space vertices pre-gathered in the
 * refine-reversed order (p1,p2,p3) = (P[vi[3j+2]], P[vi[3j+1]], P[vi[3j]]) (mesh.rs:329-331).   */
typedef struct {
  float p1[3];
  uint32_t mesh;   /* index into pbrtb200_scene.meshes */
  float p2[3];
  uint32_t attr;   /* index into tri_uv / tri_n / tri_s (per-triangle attribute records) */
  float p3[3];
  uint32_t user;   /* caller's id for this triangle (e.g. Rust prim_id); echoed, never read */
} pbrtb200_tri48;

/* One quadric: Sphere (src/shape/sphere.rs:17-25), Cylinder (src/shape/cylinder.rs:17-24) or Disk
 * (src/shape/disk.rs:15-22): world-to-object rows 0..2 of the 4x4 (affine) + params.
 *   sphere   : radius, z_min, z_max, phi_max, theta_min, theta_max
 *   cylinder : radius, z_min, z_max, phi_max                       (theta_* unused)
 *   disk     : radius, z_min = z_max = height, phi_max, theta_min = inner_radius              */
#define PBRTB200_QUADRIC_SPHERE 0u
#define PBRTB200_QUADRIC_CYLINDER 1u
#define PBRTB200_QUADRIC_DISK 2u
#define PBRTB200_QUADRIC_KIND_SHIFT 8 /* kind lives in bits 8..9 of `flip` */
typedef struct {
  float w2o[12];
  float radius, z_min, z_max, phi_max, theta_min, theta_max;
  uint32_t material;
  uint32_t flip;   /* bit 0: reverse_orientation ^ transform_swaps_handedness; bits 8..9: kind */
} pbrtb200_sphere80;

/* Per-mesh shading data (ShapeBase of the mesh, src/shape/mod.rs:33-54) */
typedef struct {
  float o2w[12];     /* rows 0..2 of object2world.m      */
  float o2w_inv[12]; /* rows 0..2 of object2world.m_inv  */
  uint32_t material;
  int32_t area_light;  /* index into lights (kind AREA) or -1 */
  uint32_t flip;       /* reverse_orientation ^ transform_swaps_handedness */
  uint32_t has_uv, has_n, has_s;
} pbrtb200_mesh;

#define PBRTB200_TEX_CONSTANT 0
#define PBRTB200_TEX_CHECKER2D 1
#define PBRTB200_TEX_UV 2
#define PBRTB200_TEX_IMAGE 3   /* ImageTexture (texture/imagemap.rs:70-73): tex1 = mipmap index */
#define PBRTB200_TEX_SCALE 4   /* ScaleTexture (texture/mod.rs:68-86): tex1 * tex2                */
#define PBRTB200_TEX_MIX 5     /* MixTexture (texture/mix.rs:9-28): tex1.lerp(tex2, tex3)          */
#define PBRTB200_TEX_BILERP 6  /* BilerpTexture (texture/bilerp.rs:10-37): value = v00,v01,v10,v11 */
#define PBRTB200_TEX_DOTS 7    /* DotsTexture (texture/dots.rs:10-47): tex1 inside, tex2 outside   */
#define PBRTB200_TEX_FBM 8     /* FBmTexture (texture/fbm.rs:8-27): value[0] = omega, aa = octaves */
#define PBRTB200_TEX_WRINKLED 9 /* WrinkledTexture (texture/fbm.rs:29-48): same fields             */
#define PBRTB200_TEX_KIND_MAX 9
#define PBRTB200_MAP_UV 0      /* UVMapping2D(su,sv,du,dv): map[0..3]                      */
#define PBRTB200_MAP_PLANAR 1  /* PlanarMapping2D(vs,vt,ds,dt): map[0..2],map[3..5],map[6..7] */
/* SphericalMapping2D / CylindricalMapping2D (texture/mapping2d.rs:106-173) and IdentityMapping3D
 * (texture/mapping3d.rs:43-66, the mapping of FBm / Wrinkled): map[0..11] = rows 0..2 of
 * world_to_texture.m (affine transforms only: the w row must be 0 0 0 1).                       */
#define PBRTB200_MAP_SPHERICAL 2
#define PBRTB200_MAP_CYLINDRICAL 3
#define PBRTB200_MAP_IDENTITY3D 4
#define PBRTB200_MAP_KIND_MAX 4
#define PBRTB200_TEX_MAX_DEPTH 3 /* textures nest (checkerboard, scale, mix, dots) to this depth */
typedef struct {
  int32_t kind;
  float value[12]; /* Constant: value[0..2] (float textures use value[0]); Bilerp: 4 x RGB;
                      FBm / Wrinkled: value[0] = omega                                         */
  int32_t map_kind;
  float map[12];
  int32_t tex1, tex2, tex3; /* children: Checkerboard (tex1, tex2), Scale, Mix (tex3 = amount), Dots */
  int32_t aa;      /* Checkerboard: 0 NONE, 1 CLOSEDFORM (texture/checkerboard.rs:10-14);
                      FBm / Wrinkled: octaves                                                  */
} pbrtb200_texture;

/* MIPMap (src/texture/mipmap.rs:143-151).  The pyramid (mipmap.rs:159-204: level 0 is the image
 * resized to powers of two, level l halves level l-1 with the 2x2 box filter) lives in
 * pbrtb200_scene.texels as RGB float4 texels, row-major per level, level l directly after level
 * l-1; level l is max(width >> l, 1) x max(height >> l, 1).  The reference stores levels in a
 * 32x32-blocked BlockedVec (utils/blocked_vec.rs) - a CPU cache layout that does not change any
 * value; on the device a row-major level puts the 8 texels of a 128-byte line side by side.
 * A float texture (TextureCache<f32>) is stored with three equal channels.                      */
#define PBRTB200_WRAP_REPEAT 0 /* ImageWrap (texture/imagewrap.rs), same order */
#define PBRTB200_WRAP_BLACK 1
#define PBRTB200_WRAP_CLAMP 2
typedef struct {
  uint32_t width, height; /* level 0 */
  uint32_t n_levels;
  uint32_t do_trilinear;
  float max_anisotropy;
  uint32_t wrap;
  uint64_t texel_offset; /* first texel of level 0 in pbrtb200_scene.texels */
} pbrtb200_mipmap;

#define PBRTB200_MAT_MATTE 0
#define PBRTB200_MAT_PLASTIC 1
typedef struct {
  int32_t kind;
  int32_t kd, sigma, ks, roughness;  /* texture indices */
  int32_t bump;  /* bump_map: Option<ScalarTextureReference> (material/matte.rs:15, plastic.rs:18):
                    0 = None (so a zero-initialised struct has no bump map), otherwise the texture
                    index of the displacement map + 1                                          */
} pbrtb200_material;

#define PBRTB200_LIGHT_POINT 0
#define PBRTB200_LIGHT_SPOT 1
#define PBRTB200_LIGHT_AREA 2 /* extension: the reference's AreaLight is a stub (area_light.rs) */
typedef struct {
  int32_t kind;
  float pos[3];
  float intensity[3];  /* point/spot: I ; area: emitted radiance L */
  float w2l[12];       /* world_to_light rows 0..2 (spot falloff) */
  float cos_total_width, cos_falloff_start;
  int32_t num_samples;
  uint32_t first_tri, n_tris; /* area: range in pbrtb200_scene.area_prims */
  float total_area;           /* filled by the library at upload */
} pbrtb200_light;

typedef struct {
  const pbrtb200_node32* nodes;
  uint32_t n_nodes;
  /* ordered primitive list (BVHAccelerator.primitives, bvh.rs:331): entry i is a triangle
   * (bit31 = 0, index into tris) or a sphere (bit31 = 1, index into spheres).               */
  const uint32_t* leaf_prim;
  uint32_t n_prims;
  const pbrtb200_tri48* tris;
  uint32_t n_tris;
  const pbrtb200_sphere80* spheres;
  const float* sphere_o2w; /* 12 floats per sphere: object2world rows 0..2 */
  uint32_t n_spheres;
  const pbrtb200_mesh* meshes;
  uint32_t n_meshes;
  const float* tri_uv; /* 6 floats per attr record (uv of p1,p2,p3) or NULL */
  const float* tri_n;  /* 9 floats per attr record or NULL */
  const float* tri_s;  /* 9 floats per attr record or NULL */
  uint32_t n_attr;
  const pbrtb200_material* materials;
  uint32_t n_materials;
  const pbrtb200_texture* textures;
  uint32_t n_textures;
  const pbrtb200_light* lights;
  uint32_t n_lights;
  /* emissive triangles of the area lights: indices into the ordered primitive list, grouped per
   * light (pbrtb200_light.first_tri / n_tris), in BVH-input (refined) order.                   */
  const uint32_t* area_prims;
  uint32_t n_area_prims;
  /* image textures: MIPMap headers + one texel pool (4 floats per texel: r, g, b, 0) */
  const pbrtb200_mipmap* mipmaps;
  uint32_t n_mipmaps;
  const float* texels;
  uint64_t n_texels;
} pbrtb200_scene;

/* ---- camera / sampler / film / integrator -------------------------------------------------- */

/* Camera::Perspective (src/camera/mod.rs:105-135, projective.rs:48-72), static camera-to-world */
typedef struct {
  float raster_to_camera[16]; /* Projection::raster_to_camera().m, row-major */
  float camera_to_world[16];
  float dx_camera[3], dy_camera[3];
  float shutter_open, shutter_close;
  float lens_radius, focal_distance;
} pbrtb200_camera;

#define PBRTB200_SAMPLER_STRATIFIED 0
#define PBRTB200_SAMPLER_LD 1
#define PBRTB200_SAMPLER_HALTON 2 /* Sampler::halton (src/sampler/halton.rs): xs = samples per pixel */
/* Sampler::stratified / low_discrepancy (src/sampler/mod.rs:30-50) over the full sample extent,
 * plus SamplerRenderer.num_tasks (sampler_renderer.rs:41-44), which fixes the per-task sub-windows
 * (sampler/base.rs:29-48) and RNG seeds (sampler_renderer.rs:74).                               */
typedef struct {
  int32_t kind;
  int32_t x_start, x_end, y_start, y_end;
  int32_t xs, ys;  /* stratified strata; LD: xs = samples per pixel (rounded up to 2^k) */
  int32_t jitter;
  float shutter_open, shutter_close;
  int32_t num_tasks;
} pbrtb200_sampler;

/* Film::Image (src/camera/film.rs:55-122) */
typedef struct {
  int32_t x_res, y_res;
  int32_t x_pixel_start, y_pixel_start, x_pixel_count, y_pixel_count;
  float filter_xw, filter_yw;
  float filter_table[256]; /* FILTER_TABLE_DIM = 16 */
} pbrtb200_film;

typedef struct {
  int32_t kind;       /* 0 = Whitted light loop (integrator/whitted.rs:30-66) */
  int32_t max_depth;  /* specular recursion contributes 0 for matte/plastic (BSDF::sample_f is
                         unimplemented in the reference), so any depth renders the same */
  int32_t strict_flags; /* 1 = reproduce BSDF::f's as-written flag test (always black) */
} pbrtb200_integrator;

/* Film pixel rectangles this call renders (multi-GPU tile partition).  A NULL tile set = the whole
 * film.  A tile set with n_rects == 0 owns NO pixel: the call renders nothing (and zeroes the film
 * unless PBRTB200_TILES_KEEP_OTHERS is set) - it is never read as "whole film". */
#define PBRTB200_TILES_KEEP_OTHERS 1u /* do not clear the film pixels outside the rects */
typedef struct {
  const int32_t* rects; /* n_rects x (x0, y0, x1, y1), half-open, in film pixel coordinates */
  uint32_t n_rects;
  uint32_t flags;       /* PBRTB200_TILES_* ; 0 = pixels outside the rects are written as zero */
} pbrtb200_tileset;

typedef struct {
  uint64_t camera_rays, camera_hits, shadow_rays;
  float ms_total;   /* device time of the whole call (CUDA events) */
  float ms_raygen, ms_trace, ms_shade, ms_shadow, ms_film;
  uint32_t kernel_launches;
  uint32_t nan_samples;
  uint32_t stack_overflows;
} pbrtb200_stats;

/* ---- wavefront records (also the unit-level parity hooks) ---------------------------------- */
typedef struct {
  float o[3], mint;
  float d[3], maxt;
} pbrtb200_ray32;
typedef struct {
  uint32_t prim; /* index into the ordered primitive list, or PBRTB200_MISS */
  float t;
  float b1, b2; /* triangle barycentrics; sphere: b1 = phi */
} pbrtb200_hit16;

/* ---- entry points -------------------------------------------------------------------------- */
int pbrtb200_create(int device, pbrtb200_ctx** out);
void pbrtb200_destroy(pbrtb200_ctx* ctx);
const char* pbrtb200_last_error(const pbrtb200_ctx* ctx); /* ctx may be NULL (create errors) */

/* Run all of this ctx's work on the caller's CUDA stream (a cudaStream_t, e.g. the host
 * framework's current stream) so that the caller's events bracket it; NULL restores the ctx's own
 * stream.  Calls stay synchronous. */
int pbrtb200_set_stream(pbrtb200_ctx* ctx, void* cuda_stream);

int pbrtb200_upload_scene(pbrtb200_ctx* ctx, const pbrtb200_scene* scene);

/* Renderer::render.  out_xyzw: 4 floats per film pixel (x_pixel_count*y_pixel_count, row-major):
 * sum(w*X), sum(w*Y), sum(w*Z), sum(w) — the contents of Film::Image.pixels (film.rs:35-41).
 * `out_is_device` != 0: out_xyzw is a device pointer on the ctx's device (no D2H copy).
 * `out_is_device` == 0: a host buffer.  Pageable memory is filled by a copy from a staging film in
 * HBM; a page-locked, mapped buffer (cudaHostAlloc, cudaHostRegister(.., cudaHostRegisterMapped))
 * is written by the film kernel directly, which is faster (the transfer overlaps the kernels).   */
int pbrtb200_render(pbrtb200_ctx* ctx, const pbrtb200_camera* cam, const pbrtb200_sampler* smp,
                    const pbrtb200_film* film, const pbrtb200_integrator* integ,
                    const pbrtb200_tileset* tiles, float* out_xyzw, int out_is_device,
                    pbrtb200_stats* stats);

/* Where does a frame's time go?  row_cost[y] (film->y_pixel_count floats, host memory) = traversal
 * cost (BVH node steps + primitive tests of one camera ray and one shadow ray per light, plus a
 * constant per sample) summed over the probe pixels of film row y, probing every stride-th pixel in
 * x and y (rows in between repeat their probe row).  Relative numbers only.  It is what
 * pbrtb200_group_render cuts row bands of equal cost with before a frame has ever been timed; the
 * reference balances by work stealing over its task queue (src/sampler_renderer.rs:168-173). */
int pbrtb200_cost_profile(pbrtb200_ctx* ctx, const pbrtb200_camera* cam, const pbrtb200_film* film, int stride,
                          float* row_cost);

/* Multi-GPU film gather without a collective (ownership of film pixels is disjoint,
 * the reference merges sub-films by pixel ownership, film.rs:149-186).  One process per GPU: the
 * gathering rank creates a film buffer and exports a CUDA IPC handle; every other rank opens it
 * and passes the mapped pointer as out_xyzw (out_is_device = 1) together with its tile set and
 * PBRTB200_TILES_KEEP_OTHERS - its film kernel then stores the owned pixels straight into the
 * gathering GPU's HBM over NVLink/NVSwitch.  The caller separates frames with a barrier.
 *   create: *dev_ptr = new buffer of n_pixels float4 on ctx's device (freed by pbrtb200_destroy),
 *           handle64 = cudaIpcMemHandle_t bytes to send to the peers
 *   open:   *dev_ptr = that buffer mapped into this process (peer access enabled lazily)
 *   close:  unmap a pointer returned by open                                                     */
int pbrtb200_peer_film_create(pbrtb200_ctx* ctx, uint64_t n_pixels, void** dev_ptr,
                              unsigned char handle64[64]);
int pbrtb200_peer_film_open(pbrtb200_ctx* ctx, const unsigned char handle64[64], void** dev_ptr);
int pbrtb200_peer_film_close(pbrtb200_ctx* ctx, void* dev_ptr);

/* ---- all the GPUs of one box behind ONE call ---------------------------------------------------
 * SamplerRenderer::render is one call that fans the frame out to every worker of the machine and
 * returns the finished film (src/sampler_renderer.rs:147-182; the pool of num_cpus threads at
 * :168-173; trait at src/renderer.rs:8-10).  A group is that for GPUs: one process, one context and
 * one host thread per device, the scene replicated on every device.  A frame is cut into contiguous
 * row bands of equal COST (first frame of a view: pbrtb200_cost_profile; later frames of the same
 * view: the bands follow each device's measured time), every device renders its band with
 * pbrtb200_render's own pipeline and
 *   - out_is_device == 0: sends ITS OWN rows straight into the caller's host film over its own PCIe
 *     link (N links in parallel, no gather, no collective).  A page-locked, mapped buffer (cudaHostAlloc,
 *     or any buffer after pbrtb200_group_pin_host_film) is written by the film kernels directly; a
 *     pageable one is filled by one staged copy per device;
 *   - out_is_device != 0: out_xyzw is a buffer on the group's FIRST device; the other devices' film
 *     kernels store their rows into it over NVLink (peer access).
 * Every film pixel is produced by exactly one device with the summation order of a single-GPU
 * render: the film is bit-identical to pbrtb200_render's for any number of devices.
 * devices == NULL: devices 0 .. n_devices-1.  stats: rays / launches summed over the devices, the
 * ms_* fields are the slowest device's.  Errors: as pbrtb200_render; pbrtb200_group_last_error names
 * the failing device.  A group call is synchronous; one call at a time per group.                  */
typedef struct pbrtb200_group pbrtb200_group;
int pbrtb200_group_create(const int* devices, int n_devices, pbrtb200_group** out);
void pbrtb200_group_destroy(pbrtb200_group* g);
const char* pbrtb200_group_last_error(const pbrtb200_group* g); /* g may be NULL (create errors) */
int pbrtb200_group_size(const pbrtb200_group* g);
pbrtb200_ctx* pbrtb200_group_ctx(pbrtb200_group* g, int i); /* device i's context (borrowed) */
int pbrtb200_group_upload_scene(pbrtb200_group* g, const pbrtb200_scene* scene);
int pbrtb200_group_render(pbrtb200_group* g, const pbrtb200_camera* cam, const pbrtb200_sampler* smp,
                          const pbrtb200_film* film, const pbrtb200_integrator* integ, float* out_xyzw,
                          int out_is_device, pbrtb200_stats* stats);
/* Page-locks and maps a host film buffer for every device of the group (cudaHostRegister), so that
 * pbrtb200_group_render's film kernels store into it directly (~25 % faster frames at 8 GPUs than staged
 * copies).  The GROUP NEVER PINS ON ITS OWN: a registration must not outlive the buffer (the old pages
 * would stay pinned and a new buffer at the same address would never receive its film), and only the
 * caller knows the buffer's lifetime.  One pinned buffer at a time (a second call replaces the first);
 * unpin — or destroy the group — BEFORE freeing the buffer.  A buffer that already is page-locked is
 * accepted as it is. */
int pbrtb200_group_pin_host_film(pbrtb200_group* g, float* xyzw, uint64_t bytes);
int pbrtb200_group_unpin_host_film(pbrtb200_group* g);
/* Device i's own stats of the last frame (rays it traced, its stage times). */
int pbrtb200_group_device_stats(const pbrtb200_group* g, int i, pbrtb200_stats* out);
/* The band cut itself (pure host arithmetic, no device needed): n_bands contiguous row bands of equal
 * summed cost over film rows y0 .. y0 + n_rows, boundaries snapped to 4 rows (sampler pixels are
 * listed in 8 x 4 tiles), non-decreasing, bounds[0] = y0, bounds[n_bands] = y0 + n_rows. */
int pbrtb200_cut_bands(const float* row_cost, int n_rows, int y0, int n_bands, int32_t* bounds);
/* The pixel work list pbrtb200_render derives from a sampler, a film and a tile set, as host arithmetic
 * (no device needed; for tools and the CPU tests): the sampler pixels whose samples can reach the film
 * pixels of the tile set (tiles == NULL: the whole film), in the order the kernels process them (8 x 4
 * tiles, row-major), each with its owning task (the reference's sub-window split) and its raster index
 * k inside that task's window; task bit 31 marks a halo pixel (outside the rects of this call).
 * Call with the arrays NULL to get *n_pixels, then with arrays of that size.  index (may be NULL):
 * one entry per pixel of the sampler extent, row-major: its list position or -1. */
int pbrtb200_work_list(const pbrtb200_sampler* smp, const pbrtb200_film* film, const pbrtb200_tileset* tiles,
                       uint32_t* n_pixels, int32_t* xy, uint32_t* k, uint32_t* task, int32_t* index);
/* The balancer pbrtb200_group_render keeps per view, as a plain host object (no device needed): for a
 * launcher that owns one process per GPU and wants the same partition, and for the CPU tests.
 * _new: bands of equal summed row_cost (NULL = uniform) over film rows y0 .. y0 + n_rows (NULL on bad
 * sizes).  _update: feed the time each band needed last frame; the boundaries follow (damped) while
 * the slowest band is more than 3 % above the mean, with a two-frame confirmation after the first 8
 * frames, at most 20 moves, and hysteresis once the bands have been within 3 % (three frames in a row
 * more than 6 % off) — a move makes every device rebuild its pixel list.  Returns 1 if the boundaries
 * moved, 0 if not, < 0 on bad arguments.  _get: bounds[0 .. n_bands]. */
typedef struct pbrtb200_bands pbrtb200_bands;
pbrtb200_bands* pbrtb200_bands_new(const float* row_cost, int n_rows, int y0, int n_bands);
void pbrtb200_bands_free(pbrtb200_bands* b);
int pbrtb200_bands_update(pbrtb200_bands* b, const float* device_ms);
int pbrtb200_bands_get(const pbrtb200_bands* b, int32_t* bounds);
/* Row bands of the last frame: bounds[0 .. n_devices] (film rows, bounds[i] .. bounds[i+1] on device i)
 * and each device's device time in ms; either pointer may be NULL. */
int pbrtb200_group_bands(const pbrtb200_group* g, int32_t* bounds, float* device_ms);

/* Scene::intersect for a batch of rays (src/scene.rs:60-63).  *_is_device: pointers are device
 * memory on the ctx's device (used by bench.py's resident-input arm).                           */
int pbrtb200_trace_closest(pbrtb200_ctx* ctx, const pbrtb200_ray32* rays, uint64_t n,
                           pbrtb200_hit16* hits, int is_device, pbrtb200_stats* stats);
/* Scene::intersect_p (src/scene.rs:65-67): occluded[i] = 1 iff any hit in [mint, maxt] */
int pbrtb200_trace_any(pbrtb200_ctx* ctx, const pbrtb200_ray32* rays, uint64_t n,
                       uint8_t* occluded, int is_device, pbrtb200_stats* stats);

/* Camera samples -> primary rays -> closest hit, one record per camera sample over the sampler
 * extent, laid out [((y - y_start) * width + (x - x_start)) * spp + i].  out_* may be NULL.
 * out_samples: 5 floats per sample (image_x, image_y, lens_u, lens_v, time).                    */
int pbrtb200_primary_hits(pbrtb200_ctx* ctx, const pbrtb200_camera* cam,
                          const pbrtb200_sampler* smp, pbrtb200_hit16* out_hits,
                          float* out_samples, pbrtb200_ray32* out_rays, int is_device,
                          pbrtb200_stats* stats);

/* HaltonSampler frames have a variable number of samples per pixel.  Per-sample outputs of
 * pbrtb200_primary_hits then use a padded layout: *cap slots per pixel of the sampler extent
 * ([((y - y_start) * width + (x - x_start)) * cap + slot]), a pixel's samples in generation order,
 * unused slots with prim = PBRTB200_MISS and NaN image coordinates.  This call returns cap (the
 * largest per-pixel count) and the number of real samples so that the caller can size them.     */
int pbrtb200_halton_layout(pbrtb200_ctx* ctx, const pbrtb200_sampler* smp, uint32_t* cap,
                           uint64_t* n_samples);

/* Film::write_image's pixel pipeline as intended (src/camera/film.rs:316-354 + write_img :15-33;
 * the code as written allocates n_pix instead of 3*n_pix floats and overwrites rgb
 * with the zero splat): per pixel  rgb = max(0, xyz_to_rgb(xyz) * (1/weight_sum)) if weight_sum
 * != 0 (spectrum.rs:31-35), then byte = (255 * rgb^(1/2.2) + 0.5).clamp(0, 255) as u8.
 * film_xyzw: n_pixels float4 (the layout pbrtb200_render produces).  out_rgb (3 floats per pixel)
 * and out_rgb8 (3 bytes per pixel) may each be NULL.  *_is_device as above; with a host film the
 * call uploads it, with host outputs only the developed image crosses PCIe (3 B/pixel).          */
int pbrtb200_film_develop(pbrtb200_ctx* ctx, const float* film_xyzw, int film_is_device,
                          uint64_t n_pixels, float* out_rgb, uint8_t* out_rgb8, int out_is_device);

/* Per-ray traversal counters of the last trace/primary call are not kept on device; the
 * algorithmic node/primitive counts used by the roofline come from the CPU oracle.              */

#ifdef __cplusplus
}
#endif
#endif /* PBRTB200_H */
