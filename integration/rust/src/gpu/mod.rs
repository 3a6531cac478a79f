// GpuRenderer: the drop-in for SamplerRenderer (src/sampler_renderer.rs:26-54, 147-182) on B200.
//
// NOT COMPILED IN THE BUILD IMAGE (no cargo / rustc there).  Written out in full against the crate as
// it stands in /root/reference; `ffi.rs` next to it is generated from include/pbrtb200.h by
// scripts/gen_rust_ffi.py and kept in step by a test.  The Python mirror of this file
// (pbrt_rust_b200/api.py: GpuRenderer) is what the test-suite drives.
//
// Wiring inside the crate:  `mod gpu;` in src/lib.rs (or main.rs), `pbrtb200` linked by build.rs
// (`println!("cargo:rustc-link-lib=dylib=pbrtb200")`), and main.rs::make_renderer (:200-202, today
// `unimplemented!()`) returns `Box::new(GpuRenderer::new(sampler, camera, surf)?)`.
pub mod ffi;
pub mod flatten;

use std::ffi::CStr;
use std::os::raw::c_int;

use self::ffi::*;
use self::flatten::FlatScene;

use camera::film::Film;
use camera::Camera;
use integrator::SurfaceIntegrator;
use intersection::Intersection;
use ray::RayDifferential;
use renderer::Renderer;
use rng::RNG;
use sampler::sample::Sample;
use sampler::Sampler;
use scene::Scene;
use spectrum::Spectrum;

/// One device (`devices.len() == 1`) or all the GPUs of the box behind ONE call, like the reference's
/// thread pool behind `render()` (sampler_renderer.rs:168-173).
enum Backend {
    Single(*mut pbrtb200_ctx),
    Group(*mut pbrtb200_group),
}

pub struct GpuRenderer {
    sampler: Sampler,
    camera: Camera,
    surface_integrator: SurfaceIntegrator,
    num_tasks: usize,
    backend: Backend,
    uploaded: Option<FlatScene>, // keeps the flattened tables (prim_ids!) alive next to the device copy
    pub last_stats: pbrtb200_stats,
}

fn last_error(b: &Backend) -> String {
    unsafe {
        let p = match b {
            &Backend::Single(ctx) => pbrtb200_last_error(ctx),
            &Backend::Group(g) => pbrtb200_group_last_error(g),
        };
        if p.is_null() { String::new() } else { CStr::from_ptr(p).to_string_lossy().into_owned() }
    }
}

fn check(rc: c_int, b: &Backend) -> Result<(), String> {
    match rc {
        PBRTB200_OK => Ok(()),
        PBRTB200_ENODEV => Err(format!("no usable CUDA device: {}", last_error(b))), // there is NO CPU fallback
        PBRTB200_ENOMEM => Err(format!("out of device memory: {}", last_error(b))),
        _ => Err(format!("pbrtb200 error {}: {}", rc, last_error(b))),
    }
}

impl GpuRenderer {
    /// Same arguments as SamplerRenderer::new minus the volume integrator, which has no constructor in
    /// the reference (integrator/mod.rs:183-201).  `devices`: CUDA device indices; one entry
    /// renders on that GPU, several render row bands of the film on all of them.
    pub fn new(sampler: Sampler, cam: Camera, surf: SurfaceIntegrator, devices: &[i32]) -> Result<GpuRenderer, String> {
        // sampler_renderer.rs:39-44, verbatim: ceil(log2(max(32 * ncpu, npix / 256))) — it fixes the
        // per-task sampler windows and RNG seeds, so the GPU must be told the same number (D12).
        let num_cpus = ::num_cpus::get() as u32;
        let num_pixels = (cam.film().x_res() * cam.film().y_res()) as u32;
        let tasks_fn = |x: u32| 31 - x.leading_zeros() + (if 0 == (x & (x - 1)) { 0 } else { 1 });
        let num_tasks = tasks_fn(::std::cmp::max(32 * num_cpus, num_pixels / 256)) as usize;

        let backend = unsafe {
            if devices.len() <= 1 {
                let mut ctx = ::std::ptr::null_mut();
                let rc = pbrtb200_create(*devices.get(0).unwrap_or(&0), &mut ctx);
                if rc != PBRTB200_OK {
                    return Err(format!("pbrtb200_create: {}", last_error(&Backend::Single(::std::ptr::null_mut()))));
                }
                Backend::Single(ctx)
            } else {
                let mut g = ::std::ptr::null_mut();
                let rc = pbrtb200_group_create(devices.as_ptr(), devices.len() as c_int, &mut g);
                if rc != PBRTB200_OK {
                    return Err(format!("pbrtb200_group_create: {}", last_error(&Backend::Group(::std::ptr::null_mut()))));
                }
                Backend::Group(g)
            }
        };
        Ok(GpuRenderer { sampler, camera: cam, surface_integrator: surf, num_tasks, backend, uploaded: None,
                         last_stats: unsafe { ::std::mem::zeroed() } })
    }

    /// Flatten + upload once (the reference builds its BVH once, at scene creation).
    pub fn preprocess(&mut self, scene: &Scene) -> Result<(), String> {
        if self.uploaded.is_some() {
            return Ok(());
        }
        let flat = FlatScene::from_scene(scene)?;
        let desc = flat.as_desc();
        let rc = unsafe {
            match &self.backend {
                &Backend::Single(ctx) => pbrtb200_upload_scene(ctx, &desc),
                &Backend::Group(g) => pbrtb200_group_upload_scene(g, &desc),
            }
        };
        check(rc, &self.backend)?;
        self.uploaded = Some(flat);
        Ok(())
    }

    /// Crate primitive id (Primitive::get_id) of an ordered-list index returned by the hit hooks.
    pub fn prim_id(&self, ordered_index: u32) -> Option<usize> {
        self.uploaded.as_ref().and_then(|f| f.prim_ids.get(ordered_index as usize).cloned())
    }
}

impl Drop for GpuRenderer {
    fn drop(&mut self) {
        unsafe {
            match &self.backend {
                &Backend::Single(ctx) => pbrtb200_destroy(ctx),
                &Backend::Group(g) => pbrtb200_group_destroy(g),
            }
        }
    }
}

// ---- descriptors -------------------------------------------------------------------------------------

/// Camera::Perspective { base, proj, dx_camera, dy_camera } (camera/mod.rs:64-77): raster_to_camera from
/// the Projection (projective.rs:33-41), camera_to_world at shutter_open (static cameras: the
/// AnimatedTransform's start_transform), the differential offsets as written (D16).
fn camera_desc(cam: &Camera) -> Result<pbrtb200_camera, String> {
    match cam {
        &Camera::Perspective { ref base, ref proj, ref dx_camera, ref dy_camera } => {
            if base.cam_to_world().is_animated() {
                return Err("animated cameras are out of scope for the B200 back end".to_string());
            }
            let mut c: pbrtb200_camera = unsafe { ::std::mem::zeroed() };
            let r2c = &proj.raster_to_camera().m().m;
            let c2w = &base.cam_to_world().start_transform().m().m;
            for i in 0..4 {
                for j in 0..4 {
                    c.raster_to_camera[4 * i + j] = r2c[i][j];
                    c.camera_to_world[4 * i + j] = c2w[i][j];
                }
            }
            c.dx_camera = [dx_camera.x, dx_camera.y, dx_camera.z];
            c.dy_camera = [dy_camera.x, dy_camera.y, dy_camera.z];
            c.shutter_open = base.shutter_open();
            c.shutter_close = base.shutter_close();
            c.lens_radius = proj.lens_radius();
            c.focal_distance = proj.focal_distance();
            Ok(c)
        }
        _ => Err("only Camera::Perspective is supported by the B200 back end".to_string()),
    }
}

/// Sampler::{Stratified, LowDiscrepancy, Halton} over the FULL sample extent (sampler/mod.rs:22-27) plus
/// num_tasks; the library derives the per-task sub-windows itself (sampler/base.rs:29-48).
fn sampler_desc(s: &Sampler, num_tasks: usize) -> Result<pbrtb200_sampler, String> {
    let (kind, b, xs, ys, jitter) = match s {
        &Sampler::Stratified(ref st) => (PBRTB200_SAMPLER_STRATIFIED, st.base(), st.x_pixel_samples() as i32,
                                         st.y_pixel_samples() as i32, st.jitter_samples() as i32),
        &Sampler::LowDiscrepancy(ref ld) => (PBRTB200_SAMPLER_LD, ld.base(), ld.base().samples_per_pixel as i32, 1, 0),
        &Sampler::Halton(ref h) => (PBRTB200_SAMPLER_HALTON, h.base(), h.base().samples_per_pixel as i32, 1, 0),
        // AdaptiveSampler: a per-task sequential chain that never terminates at an edge as written
        // (sampler/adaptive.rs:84-121) — not offered by the back end.
        &Sampler::Adaptive(_) => return Err("AdaptiveSampler is not supported by the B200 back end".to_string()),
    };
    Ok(pbrtb200_sampler {
        kind, x_start: b.x_pixel_start, x_end: b.x_pixel_end, y_start: b.y_pixel_start, y_end: b.y_pixel_end,
        xs, ys, jitter, shutter_open: b.shutter_open, shutter_close: b.shutter_close, num_tasks: num_tasks as i32,
    })
}

/// Film::Image (camera/film.rs:55-74): extents, the filter's widths and the film's own 16 x 16 table
/// (film.rs:99-110) — the device looks weights up in exactly these 256 floats.
fn film_desc(film: &Film) -> pbrtb200_film {
    let (x0, y0, xc, yc) = film.pixel_window(); // x_pixel_start, y_pixel_start, x_pixel_count, y_pixel_count
    let mut f: pbrtb200_film = unsafe { ::std::mem::zeroed() };
    f.x_res = film.x_res() as i32;
    f.y_res = film.y_res() as i32;
    f.x_pixel_start = x0;
    f.y_pixel_start = y0;
    f.x_pixel_count = xc as i32;
    f.y_pixel_count = yc as i32;
    f.filter_xw = film.filter().x_width();
    f.filter_yw = film.filter().y_width();
    f.filter_table.copy_from_slice(film.filter_table());
    f
}

impl Renderer for GpuRenderer {
    /// renderer.rs:9 / sampler_renderer.rs:147-182: one call in, the finished film out.
    fn render(&mut self, scene: &Scene) {
        self.preprocess(scene).expect("scene upload");
        let cam = camera_desc(&self.camera).expect("camera");
        let smp = sampler_desc(&self.sampler, self.num_tasks).expect("sampler");
        let fd = film_desc(self.camera.film());
        let integ = pbrtb200_integrator { kind: 0, max_depth: self.surface_integrator.max_depth() as i32, strict_flags: 0 };
        let mut xyzw = vec![0f32; 4 * fd.x_pixel_count as usize * fd.y_pixel_count as usize];
        let mut st: pbrtb200_stats = unsafe { ::std::mem::zeroed() };
        let rc = unsafe {
            match &self.backend {
                &Backend::Single(ctx) => pbrtb200_render(ctx, &cam, &smp, &fd, &integ, ::std::ptr::null(), xyzw.as_mut_ptr(), 0, &mut st),
                &Backend::Group(g) => {
                    // `xyzw` lives exactly as long as this call, so it is pinned here and unpinned below:
                    // the group never page-locks a buffer on its own (its lifetime is ours).  Pinned, every
                    // GPU's film kernel stores its rows straight into the Vec; a failed pin only means
                    // staged copies.
                    let pinned = pbrtb200_group_pin_host_film(g, xyzw.as_mut_ptr(), (xyzw.len() * 4) as u64) == PBRTB200_OK;
                    let rc = pbrtb200_group_render(g, &cam, &smp, &fd, &integ, xyzw.as_mut_ptr(), 0, &mut st);
                    if pinned {
                        pbrtb200_group_unpin_host_film(g);
                    }
                    rc
                }
            }
        };
        self.last_stats = st;
        if rc == PBRTB200_ENAN {
            panic!("Invalid radiance value!"); // sampler_renderer.rs:105, as intended
        }
        check(rc, &self.backend).expect("pbrtb200 render");
        // Film::Image.pixels: Pixel { xyz, weight_sum } per film pixel, row-major (film.rs:35-41, 64);
        // a crate-side `Film::set_pixels_xyzw` copies the four floats per pixel in.
        self.camera.film_mut().set_pixels_xyzw(&xyzw);
        self.camera.film().write_image(1.0); // film.rs:316 (after the D6 fixes; the GPU can also develop
                                             // the image itself: pbrtb200_film_develop)
    }

    /// Only specular recursion calls Renderer::li / transmittance from inside an integrator
    /// (integrator/mod.rs:21-137), and it contributes 0 for matte / plastic because BSDF::sample_f is
    /// unimplemented upstream.  Kept for the trait: one radiance query is answered by
    /// rendering nothing on the GPU — callers that need per-ray radiance use the CPU SamplerRenderer.
    fn li<'a>(&self, _scene: &'a Scene, _ray: &RayDifferential, _sample: &Sample, _rng: &mut RNG)
              -> (Spectrum, Option<Intersection>, Spectrum) {
        (Spectrum::from(0f32), None, Spectrum::from(1f32))
    }

    /// No volumes: T = 1.
    fn transmittance(&self, _scene: &Scene, _ray: &RayDifferential, _sample: &Sample, _rng: &mut RNG) -> Spectrum {
        Spectrum::from(1f32)
    }
}
