/* The drop-in boundary used from plain C, no Python and no torch: builds config 1 (the reference's
 * own 8-sphere fixture, src/primitive/aggregates/mod.rs:82-91, one point light) with the host
 * mirror, renders it through the C ABI, develops the film on the device and writes
 *   <out>.png   the 8-bit image `Film::write_image` would leave behind
 *   <out>.film  the raw film (float4 per pixel: sum w*XYZ, sum w), for the parity test
 * usage: render_c_abi <xres> <yres> <out-prefix> [--gpus N]
 * With --gpus N (N > 1) the frame goes through pbrtb200_group_render: one call in, the finished film
 * out, row bands over N devices inside the library (same film, bit for bit).
 * This is the call sequence a `GpuRenderer: Renderer` inside the Rust crate makes. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/pbrtb200.h"
#include "../include/pbrtb200_host.h"

#define CHECK(call)                                                            \
  do {                                                                         \
    int rc_ = (call);                                                          \
    if (rc_ < 0) {                                                             \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_,                      \
              grp ? pbrtb200_group_last_error(grp) : pbrtb200_last_error(ctx));  \
      return 1;                                                                \
    }                                                                          \
  } while (0)

int main(int argc, char** argv) {
  const int xres = argc > 1 ? atoi(argv[1]) : 160, yres = argc > 2 ? atoi(argv[2]) : 120;
  const char* prefix = argc > 3 ? argv[3] : "frame";
  int n_gpus = 1;
  for (int i = 4; i + 1 < argc; ++i)
    if (strcmp(argv[i], "--gpus") == 0) n_gpus = atoi(argv[i + 1]);
  pbrtb200_ctx* ctx = NULL;
  pbrtb200_group* grp = NULL;

  /* scene: Primitive::bvh(8 x Primitive::geometric(Shape::sphere(translate(v), ..), matte), 1, "sah") */
  pbh_scene* hs = pbh_scene_new();
  const float kd[3] = {0.5f, 0.5f, 0.5f}, sigma[3] = {0.f, 0.f, 0.f};
  const int mat = pbh_material_matte(hs, pbh_texture_constant(hs, kd), pbh_texture_constant(hs, sigma), -1);
  for (int i = 0; i < 8; ++i) {
    const float v[3] = {(i & 1) ? 2.f : 0.f, (i & 2) ? 2.f : 0.f, (i & 4) ? 2.f : 0.f};
    float m[16], minv[16];
    pbh_translate(v, m, minv);
    CHECK(pbh_add_sphere(hs, m, minv, 0, 1.0f, -1.0f, 1.0f, 360.0f, mat));
  }
  {
    const float lp[3] = {5.f, 6.f, -6.f}, I[3] = {50.f, 50.f, 50.f};
    float m[16], minv[16];
    pbh_translate(lp, m, minv);
    CHECK(pbh_light_point(hs, m, minv, I));
  }
  if (pbh_build_bvh(hs, 1, "sah") != 0) {
    fprintf(stderr, "pbh_build_bvh: %s\n", pbh_last_error(hs));
    return 1;
  }

  /* Camera::perspective(look_at((1,1,-6),(1,1,1),(0,1,0)).inverse(), window, 0, 0, 0, 1e6, 60, film) */
  const float pos[3] = {1.f, 1.f, -6.f}, look[3] = {1.f, 1.f, 1.f}, up[3] = {0.f, 1.f, 0.f};
  float w2c[16], c2w[16];
  CHECK(pbh_look_at(pos, look, up, w2c, c2w)); /* (m, m_inv) of the world-to-camera transform */
  const float aspect = (float)xres / (float)yres;
  float win[4];
  if (aspect > 1.f) {
    win[0] = -aspect; win[1] = aspect; win[2] = -1.f; win[3] = 1.f;
  } else {
    win[0] = -1.f; win[1] = 1.f; win[2] = -1.f / aspect; win[3] = 1.f / aspect;
  }
  pbrtb200_camera cam;
  CHECK(pbh_camera_perspective(c2w, win, 0.f, 0.f, 0.f, 1e6f, 60.f, xres, yres, &cam));
  pbrtb200_film film;
  const float crop[4] = {0.f, 1.f, 0.f, 1.f};
  CHECK(pbh_film_image(xres, yres, 0 /* box */, 0.5f, 0.5f, 0.f, 0.f, crop, &film));
  int32_t ext[4];
  pbh_film_sample_extent(&film, ext);
  pbrtb200_sampler smp;
  memset(&smp, 0, sizeof smp);
  smp.kind = PBRTB200_SAMPLER_STRATIFIED; /* Sampler::stratified(ext, 2, 2, jitter = true, 0, 0) */
  smp.x_start = ext[0]; smp.x_end = ext[1]; smp.y_start = ext[2]; smp.y_end = ext[3];
  smp.xs = 2; smp.ys = 2; smp.jitter = 1;
  smp.num_tasks = (int32_t)pbh_num_tasks(8, (uint32_t)(xres * yres)); /* SamplerRenderer::new, 8 cpus */
  pbrtb200_integrator integ;
  memset(&integ, 0, sizeof integ);
  integ.max_depth = 1;

  /* the back end */
  if (n_gpus > 1) { /* devices 0 .. N-1, one context and host thread each, scene replicated */
    if (pbrtb200_group_create(NULL, n_gpus, &grp) < 0) {
      fprintf(stderr, "pbrtb200_group_create: %s\n", pbrtb200_group_last_error(NULL));
      return 1;
    }
    CHECK(pbrtb200_group_upload_scene(grp, pbh_flat_scene(hs)));
    ctx = pbrtb200_group_ctx(grp, 0); /* borrowed: film_develop below runs on the first device */
  } else {
    CHECK(pbrtb200_create(0, &ctx));
    CHECK(pbrtb200_upload_scene(ctx, pbh_flat_scene(hs)));
  }
  const size_t npx = (size_t)film.x_pixel_count * (size_t)film.y_pixel_count;
  float* xyzw = (float*)malloc(npx * 4 * sizeof(float));
  uint8_t* rgb8 = (uint8_t*)malloc(npx * 3);
  pbrtb200_stats st;
  if (grp) {
    /* optional: page-lock the film for the group, so that every GPU's film kernel stores its rows straight
     * into it; the caller owns the buffer's lifetime, so the caller pins and unpins */
    CHECK(pbrtb200_group_pin_host_film(grp, xyzw, (uint64_t)npx * 4 * sizeof(float)));
    CHECK(pbrtb200_group_render(grp, &cam, &smp, &film, &integ, xyzw, 0, &st));
    CHECK(pbrtb200_group_unpin_host_film(grp));
  }
  else
    CHECK(pbrtb200_render(ctx, &cam, &smp, &film, &integ, NULL, xyzw, 0, &st));
  CHECK(pbrtb200_film_develop(ctx, xyzw, 0, npx, NULL, rgb8, 0));
  printf("render_c_abi: %d GPU(s), %dx%d, %llu camera rays, %llu shadow rays, %u kernel launches, %.3f ms on the device\n",
         n_gpus, film.x_pixel_count, film.y_pixel_count, (unsigned long long)st.camera_rays,
         (unsigned long long)st.shadow_rays, st.kernel_launches, st.ms_total);

  char path[1024];
  snprintf(path, sizeof path, "%s.png", prefix);
  if (pbh_write_png(path, rgb8, (uint32_t)film.x_pixel_count, (uint32_t)film.y_pixel_count) != 0) return 1;
  snprintf(path, sizeof path, "%s.film", prefix);
  FILE* f = fopen(path, "wb");
  if (!f || fwrite(xyzw, sizeof(float), npx * 4, f) != npx * 4) return 1;
  fclose(f);
  free(xyzw);
  free(rgb8);
  if (grp)
    pbrtb200_group_destroy(grp);
  else
    pbrtb200_destroy(ctx);
  pbh_scene_free(hs);
  return 0;
}
