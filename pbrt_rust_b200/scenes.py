"""Synthetic, procedurally generated scenes of BASELINE.json's named sizes.

Pure data: numpy arrays wrapped in the api.py mirror objects.  Both the CUDA product and the CPU
oracle are fed from these same descriptions.  The generator PRNG is SplitMix64 (ours; unrelated to
the renderer's StdRng)."""
import numpy as np

from .procgen import heightfield, quad_light as _quad_light, random_triangles, splitmix64  # noqa: F401
from .api import (AreaLight, Camera, CylindricalMapping2D, Film, Filter, IdentityMapping3D, Light, Material,
                  PlanarMapping2D, Primitive, SphericalMapping2D,
                  Sampler, Scene, Shape, SurfaceIntegrator, Texture, Transform, UVMapping2D)


def _matte(kd=0.5, sigma=0.0):
    return Material.matte(Texture.constant(kd), Texture.constant(sigma))


def _setup(scene, cam2world, fov, xres, yres, xs, ys, jitter, sampler="stratified", crop=(0, 1, 0, 1),
           filt=None, max_depth=1, lensr=0.0, focald=1e6):
    filt = filt or Filter.mean(0.5, 0.5)
    film = Film.image(xres, yres, filt, crop)
    aspect = xres / yres
    sw = (-aspect, aspect, -1.0, 1.0) if aspect > 1 else (-1.0, 1.0, -1.0 / aspect, 1.0 / aspect)
    cam = Camera.perspective(cam2world, sw, 0.0, 0.0, lensr, focald, fov, film)
    e = film.get_sample_extent()
    if sampler == "stratified":
        smp = Sampler.stratified(e[0], e[1], e[2], e[3], xs, ys, jitter, 0.0, 0.0)
    elif sampler == "halton":
        smp = Sampler.halton(e[0], e[1], e[2], e[3], xs * ys, 0.0, 0.0)
    else:
        smp = Sampler.low_discrepancy(e[0], e[1], e[2], e[3], xs * ys, 0.0, 0.0)
    return dict(scene=scene, camera=cam, film=film, sampler=smp, integrator=SurfaceIntegrator.whitted(max_depth))


def config1(xres=640, yres=480, sampler="stratified", crop=(0, 1, 0, 1), filt=None):
    """the reference's own 8-sphere fixture (aggregates/mod.rs:82-91), matte
    Kd 0.5, one point light I=50 at (5,6,-6), BVH sah/1, fov 60, 4 spp, Whitted depth 1."""
    mat = _matte(0.5, 0.0)
    prims = []
    for v in [(0, 0, 0), (2, 0, 0), (0, 2, 0), (2, 2, 0), (0, 0, 2), (2, 0, 2), (0, 2, 2), (2, 2, 2)]:
        t = Transform.translate(v)
        prims.append(Primitive.geometric(Shape.sphere(t, t.inverse(), False, 1.0, -1.0, 1.0, 360.0), mat))
    light = Light.point(Transform.translate((5.0, 6.0, -6.0)), 50.0)
    scene = Scene.new_with(Primitive.bvh(prims, 1, "sah"), [light])
    c2w = Transform.look_at((1, 1, -6), (1, 1, 1), (0, 1, 0)).inverse()
    return _setup(scene, c2w, 60.0, xres, yres, 2, 2, True, sampler=sampler, crop=crop, filt=filt)


def config2(n=100_000, xres=1920, yres=1080, seed=1, crop=(0, 1, 0, 1)):
    """BVH (sah/4) over n random triangles, pixel-centre samples, 1 spp."""
    vi, P = random_triangles(n, seed)
    mesh = Shape.triangle_mesh(Transform.new(), Transform.new(), False, vi, P)
    scene = Scene.new_with(Primitive.bvh([Primitive.geometric(mesh, _matte())], 4, "sah"), [])
    c2w = Transform.look_at((0, 0, -35), (0, 0, 0), (0, 1, 0)).inverse()
    return _setup(scene, c2w, 40.0, xres, yres, 1, 1, False, crop=crop)


def config3(nx=1000, nz=500, xres=1920, yres=1080, xs=4, ys=4, crop=(0, 1, 0, 1), n_lights=1,
            light_samples=1):
    """2*nx*nz-triangle heightfield (default 1 M), matte Kd 0.5, quad area
    light(s) (4x4 at y=8, L=15), Stratified xs*ys jittered, box filter, direct lighting."""
    vi, P = heightfield(nx, nz)
    mat = _matte(0.5, 0.0)
    prims = [Primitive.geometric(Shape.triangle_mesh(Transform.new(), Transform.new(), False, vi, P), mat)]
    centres = [(0.0, 0.0), (-10.0, 4.0), (10.0, -4.0), (0.0, 6.0)][:n_lights]
    for (cx, cz) in centres:
        lvi, lP = _quad_light(cx=cx, cz=cz)
        al = AreaLight(15.0, light_samples)
        prims.append(Primitive.geometric_area_light(
            Shape.triangle_mesh(Transform.new(), Transform.new(), False, lvi, lP), mat, al))
    scene = Scene.new_with(Primitive.bvh(prims, 4, "sah"), [])
    # camera chosen so that the terrain fills the frame (100 % of camera rays hit geometry)
    c2w = Transform.look_at((0, 12, -13), (0, 0, -2), (0, 1, 0)).inverse()
    return _setup(scene, c2w, 36.0, xres, yres, xs, ys, True, crop=crop)


def irradiance_probe(target=(0.0, 0.0, 0.0), emit_down=True, light_samples=1, kd=0.5, radiance=15.0,
                     res=16, spp=8, eye=None):
    """A flat Lambertian receiver (y = 0, Kd = kd) under the 4x4 quad emitter of config 3 (y = 8), seen
    through a very narrow camera aimed at `target`: every pixel sees (nearly) the same point, so the
    mean pixel value is the direct-lighting estimate there, Lo = Kd / pi * E, with E given in closed
    form by Lambert's polygon formula (`polygon_irradiance`).  emit_down=False flips the quad's
    winding: it then faces away from the receiver and E = 0.  `eye` overrides the camera position
    (e.g. below the emitter, looking up at it: the pixel value is then Le itself)."""
    gP = np.array([[-40.0, 0.0, -40.0], [40.0, 0.0, -40.0], [40.0, 0.0, 40.0], [-40.0, 0.0, 40.0]], np.float32)
    gvi = np.array([0, 1, 2, 0, 2, 3], np.uint32)
    lvi, lP = _quad_light()
    if not emit_down:
        lvi = np.array([0, 1, 2, 0, 2, 3], np.uint32)
    mat = _matte(kd, 0.0)
    prims = [Primitive.geometric(Shape.triangle_mesh(Transform.new(), Transform.new(), False, gvi, gP), mat),
             Primitive.geometric_area_light(Shape.triangle_mesh(Transform.new(), Transform.new(), False, lvi, lP),
                                            mat, AreaLight(radiance, light_samples))]
    scene = Scene.new_with(Primitive.bvh(prims, 4, "sah"), [])
    t = np.asarray(target, np.float64)
    e = np.asarray(eye, np.float64) if eye is not None else t + np.array([3.0, 5.0, -9.0])
    up = (0, 1, 0) if abs(e[0] - t[0]) + abs(e[2] - t[2]) > 1e-6 else (0, 0, 1)
    c2w = Transform.look_at(tuple(e), tuple(t), up).inverse()
    cfg = _setup(scene, c2w, 0.25, res, res, spp, spp, True)
    cfg["light_quad"] = lP.astype(np.float64)
    return cfg


def polygon_irradiance(p, n, verts, radiance):
    """Irradiance at point p (unit normal n) from a planar polygon of uniform radiance that lies
    entirely above p's horizon and is not occluded (Lambert 1760): E = L / 2 * |sum_i Gamma_i (n . g_i)|
    with Gamma_i the angle subtended by edge i and g_i the unit normal of the plane through p and edge i."""
    p, n = np.asarray(p, np.float64), np.asarray(n, np.float64)
    v = [np.asarray(q, np.float64) - p for q in verts]
    v = [q / np.linalg.norm(q) for q in v]
    acc = 0.0
    for i in range(len(v)):
        a, b = v[i], v[(i + 1) % len(v)]
        c = np.cross(a, b)
        acc += np.arccos(np.clip(np.dot(a, b), -1.0, 1.0)) * np.dot(n, c / np.linalg.norm(c))
    return radiance * abs(acc) / 2.0


def config4(n_ground=(500, 200), n_spheres=20_000, xres=3840, yres=2160, xs=8, ys=8, crop=(0, 1, 0, 1),
            seed=4):
    """textured ground (checkerboard matte, Oren-Nayar sigma 20) + jittered
    grid of plastic spheres (checker/uv Kd, Ks .25, roughness .1), one point + one area light."""
    gx, gz = n_ground
    vi, P = heightfield(gx, gz)
    uv = np.stack([(P[:, 0] + 20.0) / 40.0, (P[:, 2] + 10.0) / 20.0], -1).astype(np.float32)
    checker = Texture.checkerboard(UVMapping2D(40.0, 40.0, 0.0, 0.0), Texture.constant(0.8),
                                   Texture.constant(0.15), True)
    ground = Primitive.geometric(Shape.triangle_mesh(Transform.new(), Transform.new(), False, vi, P, uv=uv),
                                 Material.matte(checker, Texture.constant(20.0)))
    prims = [ground]
    side = max(1, int(np.ceil(np.sqrt(n_spheres * 2))))
    rows = max(1, (n_spheres + side - 1) // side)
    u = splitmix64(seed, 3 * max(n_spheres, 1)).reshape(-1, 3)
    kd_a = Texture.checkerboard(PlanarMapping2D((2, 0, 0), (0, 0, 2), 0.0, 0.0),
                                Texture.constant((0.7, 0.2, 0.2)), Texture.constant((0.2, 0.2, 0.7)), False)
    kd_b = Texture.uv(UVMapping2D(4.0, 4.0, 0.0, 0.0))
    mats = [Material.plastic(kd_a, Texture.constant(0.25), Texture.constant(0.1)),
            Material.plastic(kd_b, Texture.constant(0.25), Texture.constant(0.1))]
    for k in range(n_spheres):
        gxk, gzk = k % side, k // side
        r = 0.1 + 0.3 * u[k, 0]
        x = -19.0 + 38.0 * (gxk + 0.2 + 0.6 * u[k, 1]) / side
        z = -9.0 + 18.0 * (gzk + 0.2 + 0.6 * u[k, 2]) / rows
        t = Transform.translate((float(x), float(1.0 + r), float(z)))
        prims.append(Primitive.geometric(
            Shape.sphere(t, t.inverse(), False, float(r), float(-r), float(r), 360.0), mats[k & 1]))
    lvi, lP = _quad_light()
    prims.append(Primitive.geometric_area_light(
        Shape.triangle_mesh(Transform.new(), Transform.new(), False, lvi, lP), _matte(), AreaLight(10.0, 1)))
    scene = Scene.new_with(Primitive.bvh(prims, 4, "sah"),
                           [Light.point(Transform.translate((-8.0, 10.0, -12.0)), 120.0)])
    c2w = Transform.look_at((0, 9, -26), (0, 0, 0), (0, 1, 0)).inverse()
    return _setup(scene, c2w, 45.0, xres, yres, xs, ys, True, crop=crop)


def procedural_image(w=200, h=120, seed=11):
    """A synthetic RGB image (what read_image would return: (h, w, 3) floats = byte / 255) with
    smooth gradients, hard edges and noise, non-power-of-two on purpose (exercises the resize)."""
    y, x = np.mgrid[0:h, 0:w]
    r = (np.sin(x * 0.21) * 0.5 + 0.5) * 255
    g = ((x // 16 + y // 12) % 2) * 200 + 30
    b = (x * 255) // max(w - 1, 1)
    img = np.stack([r, g, b], -1).astype(np.int64)
    noise = (splitmix64(seed, w * h * 3).reshape(h, w, 3) * 24).astype(np.int64)
    img = np.clip(img + noise, 0, 255).astype(np.uint8)
    return img.astype(np.float32) / np.float32(255)


def textured(image=None, xres=256, yres=160, xs=2, ys=2, do_trilinear=False, max_aniso=8.0, wrap="repeat",
             gamma=2.2, nx=40, nz=20, n_spheres=24, seed=5, eye=(0, 9, -26), look=(0, 0, 0), flat=False):
    """"Next" row 2 test scene: heightfield ground with an ImageTexture Kd (UV mapping, tiled 3x2)
    and a float ImageTexture sigma, plastic spheres whose Kd is a planar-mapped image and whose
    roughness is a float image; one point light + the quad area light."""
    if image is None:
        image = procedural_image()
    vi, P = heightfield(nx, nz)
    if flat:  # no self-silhouettes: ray differentials (and with them the EWA footprints) stay sane
        P = P.copy()
        P[:, 1] = 0.0
    uv = np.stack([(P[:, 0] + 20.0) / 40.0, (P[:, 2] + 10.0) / 20.0], -1).astype(np.float32)
    kw = dict(do_trilinear=do_trilinear, max_aniso=max_aniso, wrap=wrap)
    kd_ground = Texture.image(UVMapping2D(3.0, 2.0, 0.1, 0.2), image, spectrum=True, scale=1.0, gamma=gamma, **kw)
    sig_ground = Texture.image(UVMapping2D(1.0, 1.0, 0.0, 0.0), image, spectrum=False, scale=30.0, gamma=1.0, **kw)
    prims = [Primitive.geometric(Shape.triangle_mesh(Transform.new(), Transform.new(), False, vi, P, uv=uv),
                                 Material.matte(kd_ground, sig_ground))]
    kd_sph = Texture.image(PlanarMapping2D((0.5, 0, 0), (0, 0.5, 0.25), 0.3, 0.1), image, spectrum=True,
                           scale=0.9, gamma=gamma, **kw)
    rough = Texture.image(UVMapping2D(2.0, 2.0, 0.0, 0.0), image, spectrum=False, scale=0.5, gamma=1.0, **kw)
    m_sph = Material.plastic(kd_sph, Texture.constant(0.25), rough)
    u = splitmix64(seed, 3 * n_spheres).reshape(-1, 3)
    for k in range(n_spheres):
        r = 0.4 + 0.5 * u[k, 0]
        t = Transform.translate((float(-15 + 30 * u[k, 1]), float(1.2 + r), float(-7 + 14 * u[k, 2])))
        prims.append(Primitive.geometric(Shape.sphere(t, t.inverse(), False, float(r), float(-r), float(r), 360.0), m_sph))
    lvi, lP = _quad_light()
    prims.append(Primitive.geometric_area_light(
        Shape.triangle_mesh(Transform.new(), Transform.new(), False, lvi, lP), _matte(), AreaLight(10.0, 1)))
    scene = Scene.new_with(Primitive.bvh(prims, 4, "sah"),
                           [Light.point(Transform.translate((-8.0, 10.0, -12.0)), 120.0)])
    c2w = Transform.look_at(eye, look, (0, 1, 0)).inverse()
    return _setup(scene, c2w, 45.0, xres, yres, xs, ys, True)


def quadrics(xres=160, yres=112, xs=2, ys=2, n=18, seed=9):
    """"Next" row 4 test scene: spheres, cylinders and disks (full and partial, rotated, non-uniformly
    scaled, some with reversed orientation) over a ground mesh; checker / uv / constant materials."""
    u = splitmix64(seed, 8 * n).reshape(n, 8)
    mats = [Material.matte(Texture.uv(UVMapping2D(3, 3, 0, 0)), Texture.constant(20.0)),
            Material.plastic(Texture.checkerboard(UVMapping2D(6, 6, 0, 0), Texture.constant((0.8, 0.3, 0.2)),
                                                  Texture.constant((0.2, 0.3, 0.8)), True),
                             Texture.constant(0.25), Texture.constant(0.1)),
            _matte(0.6, 0.0)]
    prims = []
    for k in range(n):
        x, z = -6.0 + 12.0 * (k % 6 + 0.5) / 6.0, -3.0 + 6.0 * (k // 6 + 0.5) / 3.0
        t = Transform.translate((float(x), 1.2 + float(u[k, 0]), float(z))) * Transform.rotate_x(float(90.0 * u[k, 1] - 60.0)) \
            * Transform.rotate_z(float(360.0 * u[k, 2])) * Transform.scale(1.0, float(0.7 + 0.6 * u[k, 3]), float(0.8 + 0.5 * u[k, 4]))
        ro = bool(k % 4 == 3)
        pm = 360.0 if k % 3 == 0 else float(120.0 + 200.0 * u[k, 5])
        if k % 3 == 0:
            shape = Shape.cylinder(t, t.inverse(), ro, float(0.3 + 0.4 * u[k, 6]), -0.7, float(0.2 + 0.6 * u[k, 7]), pm)
        elif k % 3 == 1:
            shape = Shape.disk(t, t.inverse(), ro, float(0.3 * u[k, 6]), float(0.6 + 0.4 * u[k, 7]),
                               float(0.25 * u[k, 5]), pm)
        else:
            shape = Shape.sphere(t, t.inverse(), ro, 0.7, -0.5, 0.6, pm)
        prims.append(Primitive.geometric(shape, mats[k % 3]))
    vi, P = heightfield(16, 8)
    prims.append(Primitive.geometric(Shape.triangle_mesh(Transform.new(), Transform.new(), False, vi, P), _matte(0.5, 0.0)))
    scene = Scene.new_with(Primitive.bvh(prims, 2, "sah"), [Light.point(Transform.translate((0.0, 9.0, -6.0)), 45.0)])
    c2w = Transform.look_at((0, 7, -12), (0, 0.5, 0), (0, 1, 0)).inverse()
    return _setup(scene, c2w, 45.0, xres, yres, xs, ys, True)


class TexGen:
    """Random textures / materials for the differential fuzzers.  ext=False draws exactly what the
    round-1 fuzzer drew (constant, uv, image, nested checkerboards over UV / planar mappings);
    ext=True adds spherical / cylindrical mappings, scale / mix / bilerp / dots / fbm / wrinkled
    textures and bump maps.  images=False leaves image textures out (host check of the device source)."""

    def __init__(self, rng, img, ext=False, images=True):
        self.rng, self.img, self.ext, self.images = rng, img, ext, images

    def U(self, a=0.0, b=1.0):
        return float(self.rng.uniform(a, b))

    def w2t(self):
        U = self.U
        return Transform.translate((U(-2, 2), U(-2, 2), U(-2, 2))) * Transform.rotate_x(U(0, 360)) * Transform.rotate_y(U(0, 360)) \
            * Transform.scale(U(0.3, 2.0), U(0.3, 2.0), U(0.3, 2.0))

    def mapping(self):
        rng, U = self.rng, self.U
        k = int(rng.integers(4 if self.ext else 2))
        if k == 1:
            return UVMapping2D(U(0.5, 6), U(0.5, 6), U(), U())
        if k == 0:
            return PlanarMapping2D((U(-1, 1), U(-1, 1), U(-1, 1)), (U(-1, 1), U(-1, 1), U(-1, 1)), U(), U())
        return (SphericalMapping2D if k == 2 else CylindricalMapping2D)(self.w2t() if rng.integers(3) else None)

    def image(self, spectrum, mapping, **kw):
        if not self.images:
            return Texture.constant((self.U(0.1, 0.9),) * 3 if not spectrum else (self.U(0.1, 0.9), self.U(0.1, 0.9), self.U(0.1, 0.9)))
        return Texture.image(mapping, self.img, spectrum=spectrum, do_trilinear=True, **kw)

    def noise_tex(self):
        rng, U = self.rng, self.U
        mk = Texture.fbm if rng.integers(2) else Texture.wrinkled
        return mk(int(rng.integers(0, 7)), U(0.3, 0.8), IdentityMapping3D(self.w2t() if rng.integers(3) else None))

    def spectrum_tex(self, depth=0):
        rng, U = self.rng, self.U
        if self.ext and rng.integers(2):
            k = int(rng.integers(5 if depth < 2 else 1))
            if k == 0:
                return Texture.bilerp(self.mapping(), *[(U(0, 1), U(0, 1), U(0, 1)) for _ in range(4)])
            if k == 1:
                return Texture.scale(self.spectrum_tex(depth + 1), self.unit_tex(depth + 1))
            if k == 2:
                return Texture.mix(self.spectrum_tex(depth + 1), self.spectrum_tex(depth + 1), self.unit_tex(depth + 1))
            if k == 3:
                return Texture.dots(self.mapping(), self.spectrum_tex(depth + 1), self.spectrum_tex(depth + 1))
            return Texture.scale(Texture.constant((U(0.2, 0.9), U(0.2, 0.9), U(0.2, 0.9))), self.unit_tex(depth + 1))
        k = int(rng.integers(5 if depth < 2 else 3))
        if k == 0:
            return Texture.constant((U(0.05, 0.9), U(0.05, 0.9), U(0.05, 0.9)))
        if k == 1:
            return Texture.uv(self.mapping())
        if k == 2:
            return self.image(True, self.mapping(), wrap=["repeat", "black", "clamp"][int(rng.integers(3))], scale=U(0.6, 1.0), gamma=U(1.0, 2.2))
        return Texture.checkerboard(self.mapping(), self.spectrum_tex(depth + 1), self.spectrum_tex(depth + 1), bool(rng.integers(2)))

    def unit_tex(self, depth=0):
        """a float texture with values in [0, 1] wherever u, v are in [0, 1] (every shape's are)"""
        rng, U = self.rng, self.U
        k = int(rng.integers(4 if depth < 3 else 2))
        if k == 0:
            return Texture.constant(U(0, 1))
        if k == 1:
            return Texture.bilerp(UVMapping2D(), U(0, 1), U(0, 1), U(0, 1), U(0, 1))
        if k == 2:
            return Texture.dots(self.mapping(), Texture.constant(U(0, 1)), Texture.constant(U(0, 1)))
        return Texture.checkerboard(self.mapping(), Texture.constant(U(0, 1)), Texture.constant(U(0, 1)), bool(rng.integers(2)))

    def float_tex(self, lo, hi):
        rng, U = self.rng, self.U
        if self.ext and rng.integers(2):  # lo + (hi - lo) * unit, as a mix of two constants
            return Texture.mix(Texture.constant(lo), Texture.constant(hi), self.unit_tex(1))
        if rng.integers(3) == 0:
            return self.image(False, self.mapping(), scale=hi, gamma=1.0)
        return Texture.constant(U(lo, hi))

    def bump_tex(self):
        """a displacement map of small amplitude (None half of the time)"""
        rng, U = self.rng, self.U
        k = int(rng.integers(6))
        if k >= 3:
            return None
        amp = Texture.constant(U(0.01, 0.08))
        if k == 0:
            return Texture.scale(amp, self.noise_tex())
        if k == 1:
            return Texture.scale(amp, self.unit_tex(1))
        return Texture.scale(amp, Texture.checkerboard(self.mapping(), Texture.constant(1.0), self.noise_tex(), True))

    def material(self):
        rng, U = self.rng, self.U
        if rng.integers(2):
            m = Material.matte(self.spectrum_tex(), self.float_tex(0.0, 40.0) if rng.integers(2) else Texture.constant(0.0))
        else:
            m = Material.plastic(self.spectrum_tex(), Texture.constant(U(0.05, 0.5)), self.float_tex(0.02, 0.4))
        if self.ext:
            m.bump_map = self.bump_tex()
        return m


def random_scene(seed, ext=False):
    """A seeded random small scene + camera + sampler + film, drawn from everything the back end
    supports: meshes with optional normals / tangents / uvs, spheres, cylinders, disks under random
    (also handedness-flipping) transforms, matte / plastic materials over constant, checkerboard
    (nested, antialiased or not), uv and trilinear image textures, point / spot / area lights,
    every BVH split method, every filter, stratified and LD samplers, crop windows, depth of field.
    ext=True also draws the spherical / cylindrical mappings, the scale / mix / bilerp / dots / fbm /
    wrinkled textures and bump maps (TexGen).  Used by the differential (GPU vs oracle) fuzz tests."""
    rng = np.random.default_rng(seed)
    U = lambda a=0.0, b=1.0: float(rng.uniform(a, b))
    img = procedural_image(48, 40, seed=seed + 1)
    material = TexGen(rng, img, ext=ext).material

    def xform(spread=3.0):
        t = Transform.translate((U(-spread, spread), U(0.2, 2.5), U(-spread, spread))) * Transform.rotate_x(U(0, 360)) \
            * Transform.rotate_y(U(0, 360)) * Transform.scale(U(0.5, 1.4), U(0.5, 1.4), U(0.5, 1.4) * (-1.0 if rng.integers(4) == 0 else 1.0))
        return t

    prims = []
    # ground mesh (sometimes with shading normals / tangents / uvs)
    n = int(rng.integers(4, 14))
    xs_ = np.linspace(-5, 5, n + 1)
    X, Z = np.meshgrid(xs_, xs_, indexing="ij")
    Y = U(0.0, 0.5) * np.sin(1.1 * X) * np.cos(0.8 * Z)
    P = np.stack([X, Y, Z], -1).reshape(-1, 3).astype(np.float32)
    a = (np.arange(n)[:, None] * (n + 1) + np.arange(n)[None, :]).astype(np.uint32)
    b, c, d = a + (n + 1), a + 1, a + (n + 1) + 1
    vi = np.stack([np.stack([a, c, b], -1), np.stack([b, c, d], -1)], axis=2).reshape(-1)
    N = S = UV = None
    if rng.integers(2):
        N = np.stack([rng.uniform(-0.3, 0.3, X.shape), np.ones_like(X), rng.uniform(-0.3, 0.3, X.shape)], -1).reshape(-1, 3).astype(np.float32)
    if rng.integers(2):
        S = np.stack([np.ones_like(X), rng.uniform(-0.2, 0.2, X.shape), np.zeros_like(X)], -1).reshape(-1, 3).astype(np.float32)
    if rng.integers(2):
        UV = np.stack([(X + 5) / 10, (Z + 5) / 10], -1).reshape(-1, 2).astype(np.float32)
    prims.append(Primitive.geometric(Shape.triangle_mesh(Transform.new(), Transform.new(), bool(rng.integers(2)), vi, P, N, S, UV), material()))
    for _ in range(int(rng.integers(2, 9))):
        t = xform()
        kind = int(rng.integers(4))
        ro = bool(rng.integers(2))
        pm = 360.0 if rng.integers(2) else U(60, 350)
        if kind == 0:
            shape = Shape.sphere(t, t.inverse(), ro, U(0.3, 0.9), U(-0.9, -0.2), U(0.2, 0.9), pm)
        elif kind == 1:
            shape = Shape.cylinder(t, t.inverse(), ro, U(0.2, 0.7), U(-0.8, 0.0), U(0.1, 0.9), pm)
        elif kind == 2:
            shape = Shape.disk(t, t.inverse(), ro, U(-0.3, 0.3), U(0.5, 1.0), U(0.0, 0.3), pm)
        else:
            m = int(rng.integers(1, 20))
            Pt = rng.uniform(-0.8, 0.8, (3 * m, 3)).astype(np.float32)
            shape = Shape.triangle_mesh(t, t.inverse(), ro, np.arange(3 * m, dtype=np.uint32), Pt)
        prims.append(Primitive.geometric(shape, material()))
    lights = []
    for _ in range(int(rng.integers(1, 4))):
        k = int(rng.integers(3))
        pos = (U(-5, 5), U(4, 9), U(-5, 5))
        if k == 0:
            lights.append(Light.point(Transform.translate(pos), (U(20, 90), U(20, 90), U(20, 90))))
        elif k == 1:
            l2w = Transform.look_at(pos, (U(-1, 1), 0.0, U(-1, 1)), (0, 1, 0)).inverse()
            lights.append(Light.spot(l2w, U(60, 200), U(20, 60), U(2, 15)))
        else:
            lvi, lP = _quad_light(y=U(5, 9), half=U(0.5, 2.0), cx=U(-3, 3), cz=U(-3, 3))
            prims.append(Primitive.geometric_area_light(
                Shape.triangle_mesh(Transform.new(), Transform.new(), False, lvi, lP), _matte(), AreaLight(U(4, 15), int(rng.integers(1, 4)))))
    scene = Scene.new_with(Primitive.bvh(prims, int(rng.integers(1, 5)), ["sah", "middle", "equal"][int(rng.integers(3))]), lights)
    c2w = Transform.look_at((U(-4, 4), U(3, 8), U(-10, -6)), (U(-1, 1), U(0, 1), U(-1, 1)), (0, 1, 0)).inverse()
    filt = [Filter.mean(0.5, 0.5), Filter.mean(U(0.5, 1.5), U(0.5, 1.5)), Filter.triangle(U(1, 2), U(1, 2)),
            Filter.gaussian(U(1, 2.5), U(1, 2.5), U(0.5, 2)), Filter.mitchell(2.0, 2.0, 1 / 3, 1 / 3),
            Filter.lanczos(U(1.5, 3), U(1.5, 3), 3.0)][int(rng.integers(6))]
    crop = (0, 1, 0, 1) if rng.integers(3) else tuple(sorted([U(0, 0.5), U(0.5, 1)]) + sorted([U(0, 0.5), U(0.5, 1)]))
    xres, yres = int(rng.integers(20, 72)), int(rng.integers(16, 56))
    use_ld = bool(rng.integers(3) == 0)
    lensr = U(0.02, 0.2) if rng.integers(4) == 0 else 0.0
    sampler = "ld" if use_ld else "stratified"
    if ext and rng.integers(3) == 0:  # (drawn in ext mode only: ext=False keeps the round-1 draws)
        sampler = "halton"
    return _setup(scene, c2w, U(35, 65), xres, yres, int(rng.integers(1, 4)), int(rng.integers(1, 4)), bool(rng.integers(4) != 0),
                  sampler=sampler, crop=crop, filt=filt, lensr=lensr, focald=U(6, 12))


def config5(nx=5000, nz=5000, xres=1920, yres=1080, xs=16, ys=16, crop=(0, 1, 0, 1)):
    """50 M-triangle heightfield, 4 area lights, 256 spp."""
    return config3(nx, nz, xres, yres, xs, ys, crop=crop, n_lights=4)
