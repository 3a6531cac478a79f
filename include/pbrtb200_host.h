/*
 * pbrtb200_host.h — host-side mirror of the pbrt_rust constructors that sit ABOVE the drop-in
 * boundary (Scene / Primitive::bvh / Camera::perspective / Film::image / Sampler::stratified).
 *
 * In a real integration these objects stay in the Rust crate and only the flatten shim talks to
 * pbrtb200.h.  The Rust toolchain is absent from this image, so this C/C++
 * stand-in plays the crate's part for tests and benchmarks: same constructor names, argument
 * meaning and error behaviour, and — for the BVH — the same tree, node for node, built by an
 * in-place builder (the reference moves Vecs at every level, bvh.rs:189-258).
 *
 * Everything here is plain host code (no CUDA); it never calls the CPU oracle.
 */
#ifndef PBRTB200_HOST_H
#define PBRTB200_HOST_H

#include "pbrtb200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- Transform (src/transform/transform.rs); matrices are row-major float[16] pairs (m, m_inv) */
void pbh_translate(const float v[3], float m[16], float minv[16]);
void pbh_scale(float x, float y, float z, float m[16], float minv[16]);
void pbh_rotate_x(float deg, float m[16], float minv[16]);
void pbh_rotate_y(float deg, float m[16], float minv[16]);
void pbh_rotate_z(float deg, float m[16], float minv[16]);
void pbh_mul(const float am[16], const float aminv[16], const float bm[16], const float bminv[16],
             float m[16], float minv[16]);
int pbh_invert(const float m[16], float out[16]); /* PBRTB200_ESINGULAR on "Singular matrix!" */
int pbh_look_at(const float pos[3], const float look[3], const float up[3], float m[16],
                float minv[16]);

/* ---- Scene description ---------------------------------------------------------------------- */
typedef struct pbh_scene pbh_scene;
pbh_scene* pbh_scene_new(void);
void pbh_scene_free(pbh_scene* s);
const char* pbh_last_error(const pbh_scene* s);

/* texture constructors (src/texture/ *.rs): return the texture id */
int pbh_texture_constant(pbh_scene* s, const float rgb[3]);
int pbh_texture_checkerboard(pbh_scene* s, int map_kind, const float map[12], int tex1, int tex2,
                             int antialiased);
int pbh_texture_uv(pbh_scene* s, int map_kind, const float map[12]);
/* SphericalMapping2D::new_with(xf) / CylindricalMapping2D::new_with(xf) (texture/mapping2d.rs:111-118,
 * 146-153) and IdentityMapping3D::new_with(xf) (texture/mapping3d.rs:47-55): rows 0..2 of
 * world_to_texture.m -> map[12].  PBRTB200_EINVAL when the transform is not affine.              */
int pbh_mapping_from_transform(const float m[16], float map[12]);
/* ScaleTexture::new(t1, t2) (texture/mod.rs:74-78), MixTexture::new(t1, t2, amount)
 * (texture/mix.rs:16-19), BilerpTexture::new(map, t00, t01, t10, t11) (texture/bilerp.rs:19-26; float
 * textures pass three equal channels), DotsTexture::new(mapping, inside, outside)
 * (texture/dots.rs:17-20), FBmTexture::new / WrinkledTexture::new(octaves, roughness,
 * IdentityMapping3D(world_to_texture)) (texture/fbm.rs:15-19, 36-40)                             */
int pbh_texture_scale(pbh_scene* s, int tex1, int tex2);
int pbh_texture_mix(pbh_scene* s, int tex1, int tex2, int amount);
int pbh_texture_bilerp(pbh_scene* s, int map_kind, const float map[12], const float v00[3],
                       const float v01[3], const float v10[3], const float v11[3]);
int pbh_texture_dots(pbh_scene* s, int map_kind, const float map[12], int inside, int outside);
int pbh_texture_fbm(pbh_scene* s, int octaves, float roughness, const float w2t[12]);
int pbh_texture_wrinkled(pbh_scene* s, int octaves, float roughness, const float w2t[12]);
/* TextureCache::new_texture (src/texture/imagemap.rs:128-138 / 183-193) + MIPMap::new
 * (src/texture/mipmap.rs:159-204).  rgb = what read_image returns (imagemap.rs:75-89): w*h RGB
 * texels = byte / 255, row-major from the top-left; NULL = the file could not be read, which the
 * reference turns into a 1x1 map of scale^gamma (:116-120).  spectrum != 0: TextureCache<Spectrum>
 * ((s * scale).powf(gamma) per channel); 0: TextureCache<f32> ((s.y() * scale).powf(gamma)).
 * wrap = PBRTB200_WRAP_*.  Every call builds its own MIPMap (the reference's per-file cache is a
 * memory optimisation of its loader, not part of the evaluated function).                        */
int pbh_texture_image(pbh_scene* s, int map_kind, const float map[12], const float* rgb, uint32_t w,
                      uint32_t h, int spectrum, int do_trilinear, float max_aniso, int wrap,
                      float scale, float gamma);
/* Material::matte(kd, sigma, bump_map) / Material::plastic(kd, ks, roughness, bump_map)
 * (src/material/mod.rs:88-99); bump_map = texture id of the displacement map or -1 for None      */
int pbh_material_matte(pbh_scene* s, int kd, int sigma, int bump_map);
int pbh_material_plastic(pbh_scene* s, int kd, int ks, int roughness, int bump_map);
/* PointLight::new / SpotLight::new (src/light/point.rs:21-25, spot.rs:24-35); area = extension */
int pbh_light_point(pbh_scene* s, const float l2w[16], const float l2w_inv[16], const float I[3]);
int pbh_light_spot(pbh_scene* s, const float l2w[16], const float l2w_inv[16], const float I[3],
                   float width_deg, float falloff_deg);
int pbh_light_area(pbh_scene* s, const float L[3], int num_samples);
/* Primitive::geometric(Shape::triangle_mesh(o2w, w2o, ro, vi, P, N, S, uv, None), material)
 * N, S, UV may be NULL.  area_light = id from pbh_light_area or -1.  Returns the object ordinal. */
int pbh_add_triangle_mesh(pbh_scene* s, const float o2w[16], const float o2w_inv[16], int ro,
                          const uint32_t* vi, uint64_t n_vi, const float* P, uint64_t n_p,
                          const float* N, const float* S, const float* UV, int material,
                          int area_light);
/* Primitive::geometric(Shape::cylinder(o2w, w2o, ro, rad, z0, z1, phi_max_deg), material)
 * (src/shape/cylinder.rs:27-38) and Shape::disk(o2w, w2o, ro, height, radius, inner_radius,
 * phi_max_deg) (src/shape/disk.rs:24-35).  Return the object ordinal.                          */
int pbh_add_cylinder(pbh_scene* s, const float o2w[16], const float o2w_inv[16], int ro, float rad,
                     float z0, float z1, float phi_max_deg, int material);
int pbh_add_disk(pbh_scene* s, const float o2w[16], const float o2w_inv[16], int ro, float height,
                 float radius, float inner_radius, float phi_max_deg, int material);
/* Primitive::geometric(Shape::sphere(o2w, w2o, ro, rad, z0, z1, phi_max_deg), material) */
int pbh_add_sphere(pbh_scene* s, const float o2w[16], const float o2w_inv[16], int ro, float rad,
                   float z0, float z1, float phi_max_deg, int material);
/* Primitive::bvh(prims, max_prims, sm) with sm in {"sah","middle","equal"}; unknown -> "sah" with
 * a warning, as BVHAccelerator::new does (bvh.rs:340-348).  Then flattens. */
int pbh_build_bvh(pbh_scene* s, uint32_t max_prims, const char* split_method);
/* The flatten shim's output; pointers are owned by `s` and live until it is freed / rebuilt. */
const pbrtb200_scene* pbh_flat_scene(const pbh_scene* s);
/* per ordered primitive i: (kind 0 tri / 1 sphere, source object ordinal, index within object) */
void pbh_prim_order(const pbh_scene* s, uint32_t* out3);

/* ---- Camera / Film / Sampler / Renderer ----------------------------------------------------- */
/* Camera::perspective(cam2world, screen_window, sopen, sclose, lensr, focald, fov, film) */
int pbh_camera_perspective(const float cam2world[16], const float screen_window[4], float sopen,
                           float sclose, float lensr, float focald, float fov_deg, int x_res,
                           int y_res, pbrtb200_camera* out);
/* Film::image(xres, yres, filter, crop, ..); filter_type 0 box(mean) 1 triangle 2 gaussian(p0=alpha)
 * 3 mitchell(p0=b,p1=c) 4 lanczos(p0=tau) (src/filter.rs) */
int pbh_film_image(int x_res, int y_res, int filter_type, float xw, float yw, float p0, float p1,
                   const float crop[4], pbrtb200_film* out);
void pbh_film_sample_extent(const pbrtb200_film* f, int32_t out[4]); /* film.rs:271-289 */
/* SamplerRenderer::new's task count (sampler_renderer.rs:41-44) */
uint32_t pbh_num_tasks(uint32_t num_cpus, uint32_t num_pixels);
/* Film::write_image's pixel conversion as intended (film.rs:331-346):
 * rgb = max(0, xyz_to_rgb(xyz) / weight_sum) */
void pbh_film_to_rgb(const float* xyzw, uint64_t n_pixels, float* rgb);
/* write_img's quantisation (film.rs:21-23): (255 * p^(1/2.2) + 0.5).clamp(0,255) as u8, n values */
void pbh_rgb_to_bytes(const float* rgb, uint64_t n, uint8_t* out);
/* The file `img.save(filename)` leaves behind (film.rs:15-33): 8-bit RGB PNG, row-major rgb8 */
int pbh_write_png(const char* path, const uint8_t* rgb8, uint32_t width, uint32_t height);
/* Linear float RGB as PFM (lossless companion; the crate's `exr` dependency is never called) */
int pbh_write_pfm(const char* path, const float* rgb, uint32_t width, uint32_t height);

#ifdef __cplusplus
}
#endif
#endif /* PBRTB200_HOST_H */
