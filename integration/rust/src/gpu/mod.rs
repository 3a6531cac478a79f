// Sketch of the renderer that drops in for SamplerRenderer: not compiled in the
// build image (no Rust toolchain).  `ffi.rs` next to it is generated from include/pbrtb200.h.
pub mod ffi;
pub mod flatten;
use flatten::FlatScene;

pub struct GpuRenderer { sampler: Sampler, camera: Camera, surf: SurfaceIntegrator,
                         num_tasks: usize, ctx: *mut ffi::pbrtb200_ctx, uploaded: bool }

impl GpuRenderer {
    pub fn new(sampler: Sampler, cam: Camera, surf: SurfaceIntegrator) -> Result<Self, String> {
        // same task count as SamplerRenderer::new (sampler_renderer.rs:39-44)
        let mut ctx = std::ptr::null_mut();
        check(unsafe { ffi::pbrtb200_create(0, &mut ctx) }, std::ptr::null())?;   // ENODEV: no fallback
        /* … */
    }
}

impl Renderer for GpuRenderer {
    fn render(&mut self, scene: &Scene) {
        if !self.uploaded {
            let flat = FlatScene::from_scene(scene);
            check(unsafe { ffi::pbrtb200_upload_scene(self.ctx, &flat.as_desc()) }, self.ctx).unwrap();
            self.uploaded = true;
        }
        let film = self.camera.film();
        let (cam, smp, fd) = (camera_desc(&self.camera), sampler_desc(&self.sampler, self.num_tasks),
                              film_desc(film));      // filter_table = Film's own 16x16 table
        let integ = ffi::IntegratorDesc { kind: 0, max_depth: self.surf.max_depth() as i32, strict_flags: 0 };
        let mut xyzw = vec![0f32; 4 * fd.x_pixel_count as usize * fd.y_pixel_count as usize];
        let mut st = ffi::Stats::default();
        let rc = unsafe { ffi::pbrtb200_render(self.ctx, &cam, &smp, &fd, &integ, std::ptr::null(),
                                                xyzw.as_mut_ptr(), 0, &mut st) };
        if rc == -3 { panic!("Invalid radiance value!"); }       // sampler_renderer.rs:105 intent
        check(rc, self.ctx).unwrap();
        self.camera.film_mut().set_pixels_xyzw(&xyzw);           // Pixel{xyz, weight_sum}, film.rs:35-41
        self.camera.film().write_image(1.0);
    }
    // li / transmittance are only used for specular recursion, which the GPU path handles
    // internally (and which contributes 0 for matte/plastic): delegate to a CPU SamplerRenderer.
    fn li(&self, /* … */) -> (Spectrum, Option<Intersection>, Spectrum) { unimplemented!() }
    fn transmittance(&self, /* … */) -> Spectrum { Spectrum::from(1f32) }
}
