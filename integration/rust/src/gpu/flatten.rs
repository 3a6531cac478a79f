// Scene flattening for the B200 back end: BVHAccelerator + primitives + materials + textures + lights
// -> the POD tables of include/pbrtb200.h (bound in ffi.rs).
//
// NOT COMPILED IN THE BUILD IMAGE: there is no Rust toolchain there (no cargo / rustc).  The file is
// written out in full against the crate as it stands in /root/reference; the C++ host mirror
// (pbrt_rust_b200/csrc/host/pbrt_host.cpp, `pbh_build`) is the same algorithm, compiled and tested,
// and every step below names the lines of it that it corresponds to.
//
// The module lives INSIDE the crate (src/gpu/) because nearly every field it reads is private.  The
// crate-side additions it needs are all `pub(crate)` accessors, listed once here and again in
// INTEGRATION.md §2:
//   primitive/aggregates/bvh.rs   BVHAccelerator::{nodes(), primitives()}            (fields :329-332)
//   primitive/mod.rs              Primitive::{as_geometric(), as_bvh()}              (enum Prim :66-70)
//   primitive/geometric.rs        GeometricPrimitive::{shape(), material()}          (fields :19-23)
//   shape/mesh.rs                 Triangle::{mesh(), v()}, Mesh::{p(), n(), s(), uvs()}  (:27-30, :277-285)
//   shape/{sphere,cylinder,disk}.rs   one `params()` each returning the private scalars   (:17-25, :17-24, :15-22)
//   transform/transform.rs        Transform::{m(), m_inv()}                          (fields :15-19)
//   scene.rs                      Scene::aggregate()                                 (field :16)
//   material/{matte,plastic}.rs   field accessors                                    (:13-17, :15-20)
//   texture/mod.rs                one method on the Texture trait, `flatten` (below)  (:36-46)
//   texture/mapping2d.rs, mapping3d.rs   `flatten_mapping` on the two mapping traits  (:36-47, :30-41)
//   light/mod.rs                  one method on the Light trait, `flatten` (below)    (:49-52)
//   area_light.rs                 AreaLight { l_emit: Spectrum, num_samples: usize } — the reference's
//                                 AreaLight is a unit struct (:9-24); the oracle-defined diffuse area
//                                 light needs these two fields.
use std::collections::HashMap;
use std::sync::Arc;

use super::ffi::*;

use area_light::AreaLight;
use light::Light;
use material::Material;
use primitive::aggregates::bvh::{BVHAccelerator, PackedBVHNode};
use primitive::Primitive;
use scene::Scene;
use shape::Shape;
use spectrum::Spectrum;
use texture::mipmap::MIPMap;
use texture::Texture;
use transform::transform::Transform;

pub const LEAF_PRIM_QUADRIC: u32 = 0x8000_0000;

/// Rows 0..2 of a 4x4 (the device only stores affine transforms; row 3 is 0 0 0 1).
fn rows3(m: &[[f32; 4]; 4]) -> [f32; 12] {
    let mut r = [0f32; 12];
    for i in 0..3 {
        for j in 0..4 {
            r[4 * i + j] = m[i][j];
        }
    }
    r
}

fn rgb(s: &Spectrum) -> [f32; 3] {
    s.to_rgb() // spectrum.rs:472 — the back end computes in RGB, like Spectrum::RGB
}

// ---------------------------------------------------------------------------------------------------
// Textures.  `Arc<dyn Texture<T>>` trait objects cannot be matched on from outside, so the crate's own
// trait gets one more method (texture/mod.rs:36-46):
//
//     pub trait Texture<T>: Debug + Send + Sync + internal::TextureBase<T> {
//         fn evaluate(&self, _: &DifferentialGeometry) -> T;
//         fn flatten(&self, table: &mut TexTable) -> i32;      // <- new: push own record, return its index
//     }
//
// implemented per texture type exactly as `impl FlattenTexture for ...` below (the blanket impl at
// texture/mod.rs:40-46 forwards to TextureBase, where the per-type bodies go).  Children are flattened
// first, so a parent's tex1 / tex2 / tex3 are indices that already exist; records are deduplicated by
// Arc pointer so a texture shared by many materials is stored once.
// Host mirror: pbh_texture_* (pbrt_host.cpp), one call per constructor.
// ---------------------------------------------------------------------------------------------------
pub struct TexTable {
    pub textures: Vec<pbrtb200_texture>,
    pub mipmaps: Vec<pbrtb200_mipmap>,
    pub texels: Vec<f32>, // 4 floats per texel (r, g, b, 0), levels back to back
    seen: HashMap<usize, i32>,      // Arc data pointer -> texture index
    seen_mips: HashMap<usize, i32>, // Arc<MIPMap> pointer -> mipmap index
}

impl TexTable {
    pub fn new() -> TexTable {
        TexTable { textures: vec![], mipmaps: vec![], texels: vec![], seen: HashMap::new(), seen_mips: HashMap::new() }
    }

    fn blank(kind: i32) -> pbrtb200_texture {
        pbrtb200_texture { kind, value: [0.0; 12], map_kind: PBRTB200_MAP_UV, map: [0.0; 12], tex1: -1, tex2: -1, tex3: -1, aa: 0 }
    }

    fn push(&mut self, t: pbrtb200_texture) -> i32 {
        self.textures.push(t);
        (self.textures.len() - 1) as i32
    }

    /// Entry point used by materials: flatten `t` once per distinct Arc.
    pub fn add<T>(&mut self, t: &Arc<dyn Texture<T>>) -> i32 {
        let key = Arc::as_ptr(t) as *const () as usize;
        if let Some(&i) = self.seen.get(&key) {
            return i;
        }
        let i = t.flatten(self);
        self.seen.insert(key, i);
        i
    }

    /// MIPMap (texture/mipmap.rs:143-151) -> header + texels.  The pyramid is copied level by level,
    /// row-major (BlockedVec::get(s, t), utils/blocked_vec.rs), 3 equal channels for f32 maps.
    pub fn add_mipmap<M: MipTexel>(&mut self, m: &Arc<MIPMap<M>>) -> i32 {
        let key = Arc::as_ptr(m) as usize;
        if let Some(&i) = self.seen_mips.get(&key) {
            return i;
        }
        let hdr = pbrtb200_mipmap {
            width: m.width() as u32,
            height: m.height() as u32,
            n_levels: m.levels() as u32,
            do_trilinear: m.do_trilinear() as u32,
            max_anisotropy: m.max_anisotropy(),
            wrap: m.wrap_mode() as u32, // ImageWrap order == PBRTB200_WRAP_* order (texture/imagewrap.rs)
            texel_offset: (self.texels.len() / 4) as u64,
        };
        for l in 0..m.levels() {
            let (w, h) = (std::cmp::max(m.width() >> l, 1), std::cmp::max(m.height() >> l, 1));
            for t in 0..h {
                for s in 0..w {
                    let c = m.level(l).get(s, t).to_rgb3();
                    self.texels.extend_from_slice(&[c[0], c[1], c[2], 0.0]);
                }
            }
        }
        self.mipmaps.push(hdr);
        let i = (self.mipmaps.len() - 1) as i32;
        self.seen_mips.insert(key, i);
        i
    }
}

/// What a MIPMap texel contributes to the RGB texel pool (Spectrum -> rgb, f32 -> (v, v, v)).
pub trait MipTexel: Default + Clone {
    fn to_rgb3(&self) -> [f32; 3];
}
impl MipTexel for f32 {
    fn to_rgb3(&self) -> [f32; 3] { [*self, *self, *self] }
}
impl MipTexel for Spectrum {
    fn to_rgb3(&self) -> [f32; 3] { self.to_rgb() }
}

/// A texture value as the 3 floats of `pbrtb200_texture.value` (float textures use value[0]).
pub trait TexValue: Clone {
    fn to_value3(&self) -> [f32; 3];
}
impl TexValue for f32 {
    fn to_value3(&self) -> [f32; 3] { [*self, *self, *self] }
}
impl TexValue for Spectrum {
    fn to_value3(&self) -> [f32; 3] { self.to_rgb() }
}

/// (map_kind, map[12]) of a mapping; the two mapping traits get `fn flatten_mapping(&self) -> (i32, [f32; 12])`.
pub fn uv_mapping(su: f32, sv: f32, du: f32, dv: f32) -> (i32, [f32; 12]) {
    // UVMapping2D { su, sv, du, dv } (texture/mapping2d.rs:49-54)
    let mut m = [0f32; 12];
    m[0] = su; m[1] = sv; m[2] = du; m[3] = dv;
    (PBRTB200_MAP_UV, m)
}
pub fn planar_mapping(vs: [f32; 3], vt: [f32; 3], ds: f32, dt: f32) -> (i32, [f32; 12]) {
    // PlanarMapping2D { vs, vt, ds, dt } (texture/mapping2d.rs:175-180)
    let mut m = [0f32; 12];
    m[0..3].copy_from_slice(&vs);
    m[3..6].copy_from_slice(&vt);
    m[6] = ds; m[7] = dt;
    (PBRTB200_MAP_PLANAR, m)
}
pub fn world_to_texture_mapping(kind: i32, w2t: &Transform) -> (i32, [f32; 12]) {
    // SphericalMapping2D / CylindricalMapping2D / IdentityMapping3D { world_to_texture }
    // (texture/mapping2d.rs:107-109, 142-144; mapping3d.rs:43-45): rows 0..2 of world_to_texture.m
    (kind, rows3(&w2t.m().m))
}

// The per-type bodies of Texture::flatten.  Field names are the crate's (cited per impl).
pub trait FlattenTexture {
    fn flatten_into(&self, table: &mut TexTable) -> i32;
}

// ConstantTexture { value } (texture/mod.rs:52-54)
impl<T: TexValue> FlattenTexture for texture::ConstantTexture<T> {
    fn flatten_into(&self, table: &mut TexTable) -> i32 {
        let mut t = TexTable::blank(PBRTB200_TEX_CONSTANT);
        t.value[0..3].copy_from_slice(&self.value().to_value3());
        table.push(t)
    }
}

// ScaleTexture { tex1, tex2 } (texture/mod.rs:68-86): tex1 * tex2
impl<T1, T2> FlattenTexture for texture::ScaleTexture<T1, T2> {
    fn flatten_into(&self, table: &mut TexTable) -> i32 {
        let (a, b) = (table.add(self.tex1()), table.add(self.tex2()));
        let mut t = TexTable::blank(PBRTB200_TEX_SCALE);
        t.tex1 = a; t.tex2 = b;
        table.push(t)
    }
}

// CheckerboardTexture { mapping, tex1, tex2, aa_method } (texture/checkerboard.rs:17-22)
impl<T> FlattenTexture for texture::checkerboard::CheckerboardTexture<T> {
    fn flatten_into(&self, table: &mut TexTable) -> i32 {
        let (a, b) = (table.add(self.tex1()), table.add(self.tex2()));
        let (mk, m) = self.mapping().flatten_mapping();
        let mut t = TexTable::blank(PBRTB200_TEX_CHECKER2D);
        t.map_kind = mk; t.map = m; t.tex1 = a; t.tex2 = b;
        t.aa = self.aa_is_closed_form() as i32; // CheckerboardAA::{NONE = 0, CLOSEDFORM = 1} (:10-14)
        table.push(t)
    }
}

// UVTexture { mapping } (texture/uv.rs:10-12)
impl FlattenTexture for texture::uv::UVTexture {
    fn flatten_into(&self, table: &mut TexTable) -> i32 {
        let (mk, m) = self.mapping().flatten_mapping();
        let mut t = TexTable::blank(PBRTB200_TEX_UV);
        t.map_kind = mk; t.map = m;
        table.push(t)
    }
}

// MixTexture { tex1, tex2, amount } (texture/mix.rs:9-13): tex1.lerp(tex2, amount)
impl<T> FlattenTexture for texture::mix::MixTexture<T> where T: ::utils::Lerp<f32> {
    fn flatten_into(&self, table: &mut TexTable) -> i32 {
        let (a, b, c) = (table.add(self.tex1()), table.add(self.tex2()), table.add(self.amount()));
        let mut t = TexTable::blank(PBRTB200_TEX_MIX);
        t.tex1 = a; t.tex2 = b; t.tex3 = c;
        table.push(t)
    }
}

// BilerpTexture { mapping, v00, v01, v10, v11 } (texture/bilerp.rs:10-16)
impl<T: TexValue> FlattenTexture for texture::bilerp::BilerpTexture<T> where T: ::utils::Lerp<f32> {
    fn flatten_into(&self, table: &mut TexTable) -> i32 {
        let (mk, m) = self.mapping().flatten_mapping();
        let mut t = TexTable::blank(PBRTB200_TEX_BILERP);
        t.map_kind = mk; t.map = m;
        for (k, v) in [self.v00(), self.v01(), self.v10(), self.v11()].iter().enumerate() {
            t.value[3 * k..3 * k + 3].copy_from_slice(&v.to_value3());
        }
        table.push(t)
    }
}

// DotsTexture { mapping, inside_dot, outside_dot } (texture/dots.rs:10-14)
impl<T> FlattenTexture for texture::dots::DotsTexture<T> {
    fn flatten_into(&self, table: &mut TexTable) -> i32 {
        let (a, b) = (table.add(self.inside_dot()), table.add(self.outside_dot()));
        let (mk, m) = self.mapping().flatten_mapping();
        let mut t = TexTable::blank(PBRTB200_TEX_DOTS);
        t.map_kind = mk; t.map = m; t.tex1 = a; t.tex2 = b;
        table.push(t)
    }
}

// FBmTexture / WrinkledTexture { omega, octaves, mapping } (texture/fbm.rs:9-13, 29-33)
impl FlattenTexture for texture::fbm::FBmTexture {
    fn flatten_into(&self, table: &mut TexTable) -> i32 {
        let (mk, m) = self.mapping().flatten_mapping(); // IdentityMapping3D -> PBRTB200_MAP_IDENTITY3D
        let mut t = TexTable::blank(PBRTB200_TEX_FBM);
        t.map_kind = mk; t.map = m; t.value[0] = self.omega(); t.aa = self.octaves();
        table.push(t)
    }
}
impl FlattenTexture for texture::fbm::WrinkledTexture {
    fn flatten_into(&self, table: &mut TexTable) -> i32 {
        let (mk, m) = self.mapping().flatten_mapping();
        let mut t = TexTable::blank(PBRTB200_TEX_WRINKLED);
        t.map_kind = mk; t.map = m; t.value[0] = self.omega(); t.aa = self.octaves();
        table.push(t)
    }
}

// ImageTexture { mipmap, mapping } (texture/imagemap.rs:70-73)
impl<M: MipTexel> FlattenTexture for texture::imagemap::ImageTexture<M> {
    fn flatten_into(&self, table: &mut TexTable) -> i32 {
        let mip = table.add_mipmap(self.mipmap());
        let (mk, m) = self.mapping().flatten_mapping();
        let mut t = TexTable::blank(PBRTB200_TEX_IMAGE);
        t.map_kind = mk; t.map = m; t.tex1 = mip;
        table.push(t)
    }
}

// ---------------------------------------------------------------------------------------------------
// Lights.  `Arc<dyn Light>` again: one method on the trait (light/mod.rs:49-52),
//     fn flatten(&self) -> pbrtb200_light;
// with the bodies below.  Area lights are not in Scene::lights in the reference (AreaLight is a stub);
// the shim appends one record per distinct Arc<AreaLight> found on the primitives.
// Host mirror: pbh_light_point / pbh_light_spot / pbh_light_area.
// ---------------------------------------------------------------------------------------------------
fn blank_light(kind: i32) -> pbrtb200_light {
    pbrtb200_light { kind, pos: [0.0; 3], intensity: [0.0; 3], w2l: [0.0; 12], cos_total_width: 0.0, cos_falloff_start: 0.0,
                     num_samples: 1, first_tri: 0, n_tris: 0, total_area: 0.0 }
}

// PointLight { base, light_pos, intensity } (light/point.rs:14-18)
pub fn flatten_point_light(light_pos: [f32; 3], intensity: &Spectrum) -> pbrtb200_light {
    let mut l = blank_light(PBRTB200_LIGHT_POINT);
    l.pos = light_pos;
    l.intensity = rgb(intensity);
    l
}

// SpotLight { base, light_pos, intensity, cos_total_width, cos_falloff_start } (light/spot.rs:16-22);
// base.world_to_light (light/mod.rs:17-21) carries the cone axis (spot.rs:46: world_to_light.xf(-w)).
pub fn flatten_spot_light(light_pos: [f32; 3], intensity: &Spectrum, world_to_light: &Transform, cos_total_width: f32,
                          cos_falloff_start: f32) -> pbrtb200_light {
    let mut l = blank_light(PBRTB200_LIGHT_SPOT);
    l.pos = light_pos;
    l.intensity = rgb(intensity);
    l.w2l = rows3(&world_to_light.m().m);
    l.cos_total_width = cos_total_width;
    l.cos_falloff_start = cos_falloff_start;
    l
}

pub fn flatten_area_light(al: &AreaLight) -> pbrtb200_light {
    let mut l = blank_light(PBRTB200_LIGHT_AREA);
    l.intensity = rgb(al.l_emit());           // emitted radiance L
    l.num_samples = al.num_samples() as i32;  // shadow rays per camera sample
    l                                         // first_tri / n_tris are filled below, total_area by the library
}

// ---------------------------------------------------------------------------------------------------
// The flattened scene.  Owns every Vec the pbrtb200_scene descriptor points into.
// ---------------------------------------------------------------------------------------------------
pub struct FlatScene {
    pub nodes: Vec<pbrtb200_node32>,
    pub leaf_prim: Vec<u32>,
    pub tris: Vec<pbrtb200_tri48>,
    pub spheres: Vec<pbrtb200_sphere80>,
    pub sphere_o2w: Vec<f32>,
    pub meshes: Vec<pbrtb200_mesh>,
    pub tri_uv: Vec<f32>,
    pub tri_n: Vec<f32>,
    pub tri_s: Vec<f32>,
    pub materials: Vec<pbrtb200_material>,
    pub tex: TexTable,
    pub lights: Vec<pbrtb200_light>,
    pub area_prims: Vec<u32>,
    /// ordered-primitive index -> Primitive::get_id() (primitive/mod.rs:125): results of
    /// pbrtb200_primary_hits / trace_closest are reported in the crate's own ids through this table
    /// (SURVEY D19 / Appendix A).
    pub prim_ids: Vec<usize>,
    has_quadrics: bool,
    any_uv: bool,
    any_n: bool,
    any_s: bool,
}

impl FlatScene {
    pub fn from_scene(scene: &Scene) -> Result<FlatScene, String> {
        // Scene.aggregate must be Primitive::bvh(..) (primitive/mod.rs:112-117): the back end traverses
        // the reference's own tree.  Grid / KdTree aggregates are out of scope.
        let bvh: &BVHAccelerator = scene.aggregate().as_bvh().ok_or("the B200 back end needs a BVH aggregate")?;
        let mut fs = FlatScene {
            nodes: Vec::with_capacity(bvh.nodes().len()), leaf_prim: vec![], tris: vec![], spheres: vec![],
            sphere_o2w: vec![], meshes: vec![], tri_uv: vec![], tri_n: vec![], tri_s: vec![], materials: vec![],
            tex: TexTable::new(), lights: vec![], area_prims: vec![], prim_ids: vec![], has_quadrics: false,
            any_uv: false, any_n: false, any_s: false,
        };

        // 1. Nodes: a straight copy of BVHAccelerator.nodes, already in depth-first order with the first
        //    child at i + 1 (bvh.rs:260-327).  [host mirror: Builder::build writes the same array]
        for n in bvh.nodes() {
            fs.nodes.push(match n {
                &PackedBVHNode::Leaf { ref bounds, prim_offset, num_prims } => pbrtb200_node32 {
                    bmin: [bounds.p_min.x, bounds.p_min.y, bounds.p_min.z],
                    bmax: [bounds.p_max.x, bounds.p_max.y, bounds.p_max.z],
                    offset: prim_offset as u32, count: num_prims as u16, axis: 0, is_leaf: 1,
                },
                &PackedBVHNode::Inner { ref bounds, second_child_offset, axis } => pbrtb200_node32 {
                    bmin: [bounds.p_min.x, bounds.p_min.y, bounds.p_min.z],
                    bmax: [bounds.p_max.x, bounds.p_max.y, bounds.p_max.z],
                    offset: second_child_offset as u32, count: 0, axis: axis as u8, is_leaf: 0,
                },
            });
        }

        // 2. Lights of the scene, in Scene::lights order (the Whitted loop iterates them in that order,
        //    whitted.rs:49).  Area lights follow, one per distinct Arc<AreaLight>, in the order their first
        //    primitive appears in the ordered list.
        for l in scene.lights() {
            fs.lights.push(l.flatten());
        }
        let prims: &Vec<Primitive> = bvh.primitives(); // bvh.rs:331, already in leaf order
        let mut area_index: HashMap<usize, i32> = HashMap::new();
        let mut area_tris: Vec<Vec<u32>> = vec![];
        for p in prims.iter() {
            if let Some(al) = p.as_geometric().and_then(|g| g.area_light()) {
                let key = Arc::as_ptr(&al) as usize;
                if !area_index.contains_key(&key) {
                    area_index.insert(key, fs.lights.len() as i32);
                    fs.lights.push(flatten_area_light(&al));
                    area_tris.push(vec![]);
                }
            }
        }
        let first_area = fs.lights.len() - area_tris.len();

        // 3. Materials, deduplicated by Arc pointer (many primitives share one material).
        let mut mat_index: HashMap<usize, u32> = HashMap::new();
        // 4. Meshes, deduplicated by Arc<Mesh> pointer: one pbrtb200_mesh per (mesh, material, area light).
        let mut mesh_index: HashMap<(usize, u32, i32), u32> = HashMap::new();

        // A first pass decides which per-triangle attribute arrays exist at all (a scene without uvs /
        // normals / tangents does not upload them).  [host mirror: any_uv / any_n / any_s]
        for p in prims.iter() {
            if let Some(&Shape::Triangle(ref t)) = p.as_geometric().map(|g| g.shape()) {
                fs.any_uv |= t.mesh().uvs().is_some();
                fs.any_n |= t.mesh().n().is_some();
                fs.any_s |= t.mesh().s().is_some();
            } else {
                fs.has_quadrics = true;
            }
        }

        // 5. The ordered primitive list.
        for (i, p) in prims.iter().enumerate() {
            let g = p.as_geometric().ok_or("only geometric primitives can be leaves of the flattened BVH")?;
            let material = {
                let key = Arc::as_ptr(g.material()) as usize;
                match mat_index.get(&key) {
                    Some(&m) => m,
                    None => {
                        let m = fs.flatten_material(g.material())?;
                        mat_index.insert(key, m);
                        m
                    }
                }
            };
            let area_light = g.area_light().map(|al| area_index[&(Arc::as_ptr(&al) as usize)]).unwrap_or(-1);
            fs.prim_ids.push(p.get_id());
            match g.shape() {
                &Shape::Triangle(ref t) => {
                    let mesh = t.mesh();
                    let base = mesh.base();
                    let mkey = (Arc::as_ptr(mesh) as usize, material, area_light);
                    let mi = match mesh_index.get(&mkey) {
                        Some(&m) => m,
                        None => {
                            fs.meshes.push(pbrtb200_mesh {
                                o2w: rows3(&base.object2world.m().m),
                                o2w_inv: rows3(&base.object2world.m_inv().m),
                                material, area_light,
                                flip: (base.reverse_orientation ^ base.transform_swaps_handedness) as u32, // shape/mod.rs:33-54
                                has_uv: mesh.uvs().is_some() as u32,
                                has_n: mesh.n().is_some() as u32,
                                has_s: mesh.s().is_some() as u32,
                            });
                            let m = (fs.meshes.len() - 1) as u32;
                            mesh_index.insert(mkey, m);
                            m
                        }
                    };
                    // Triangle.v is already the refine-reversed order (mesh.rs:329-331: indices are
                    // popped from the back); Mesh.p is in world space (mesh.rs:300).
                    let v = t.v();
                    let p3 = |k: usize| { let q = &mesh.p()[v[k]]; [q.x, q.y, q.z] };
                    let ti = fs.tris.len() as u32;
                    fs.tris.push(pbrtb200_tri48 { p1: p3(0), mesh: mi, p2: p3(1), attr: ti, p3: p3(2), user: p.get_id() as u32 });
                    if fs.has_quadrics {
                        fs.leaf_prim.push(ti);
                    }
                    if fs.any_uv {
                        for k in 0..3 {
                            match mesh.uvs() {
                                Some(uv) => { fs.tri_uv.push(uv[2 * v[k]]); fs.tri_uv.push(uv[2 * v[k] + 1]); }
                                None => { fs.tri_uv.push(0.0); fs.tri_uv.push(0.0); }
                            }
                        }
                    }
                    if fs.any_n {
                        for k in 0..3 {
                            match mesh.n() {
                                Some(n) => fs.tri_n.extend_from_slice(&[n[v[k]].x, n[v[k]].y, n[v[k]].z]),
                                None => fs.tri_n.extend_from_slice(&[0.0; 3]),
                            }
                        }
                    }
                    if fs.any_s {
                        for k in 0..3 {
                            match mesh.s() {
                                Some(s) => fs.tri_s.extend_from_slice(&[s[v[k]].x, s[v[k]].y, s[v[k]].z]),
                                None => fs.tri_s.extend_from_slice(&[0.0; 3]),
                            }
                        }
                    }
                    if area_light >= 0 {
                        area_tris[area_light as usize - first_area].push(i as u32);
                    }
                }
                q @ &Shape::Sphere(_) | q @ &Shape::Cylinder(_) | q @ &Shape::Disk(_) => {
                    if area_light >= 0 {
                        return Err("area lights are supported on triangle meshes only".to_string());
                    }
                    let si = fs.spheres.len() as u32;
                    fs.push_quadric(q, material);
                    fs.leaf_prim.push(LEAF_PRIM_QUADRIC | si);
                }
                // BVHAccelerator::new fully refines its input (bvh.rs:337-345), so meshes never reach
                // the ordered list unrefined; LoopSubdiv refines to a mesh (loopsubdiv.rs) first.
                _ => return Err("unrefined shape in the ordered primitive list".to_string()),
            }
        }

        // 6. Emissive triangles per area light: indices into the ordered list.  The reference order of
        //    a light's triangles is the BVH-INPUT (refined) order, which prim ids preserve
        //    (NEXT_PRIM_ID is monotonic, primitive/mod.rs:31-36).  [host mirror: pos_of_ref]
        for (k, tris) in area_tris.iter_mut().enumerate() {
            if tris.is_empty() {
                return Err("area light without emissive triangles".to_string());
            }
            tris.sort_by_key(|&i| fs.prim_ids[i as usize]);
            let l = &mut fs.lights[first_area + k];
            l.first_tri = fs.area_prims.len() as u32;
            l.n_tris = tris.len() as u32;
            fs.area_prims.extend_from_slice(tris);
        }
        Ok(fs)
    }

    // MatteMaterial { sigma, bump_map, k_d } (material/matte.rs:13-17),
    // PlasticMaterial { k_d, k_s, roughness, bump_map } (material/plastic.rs:15-20).
    fn flatten_material(&mut self, m: &Arc<Material>) -> Result<u32, String> {
        let rec = match m.as_ref() {
            &Material::Matte(ref mm) => pbrtb200_material {
                kind: PBRTB200_MAT_MATTE,
                kd: self.tex.add(mm.k_d()),
                sigma: self.tex.add(mm.sigma()),
                ks: 0, roughness: 0,
                bump: mm.bump_map().map(|b| self.tex.add(b) + 1).unwrap_or(0),
            },
            &Material::Plastic(ref pm) => pbrtb200_material {
                kind: PBRTB200_MAT_PLASTIC,
                kd: self.tex.add(pm.k_d()),
                sigma: 0,
                ks: self.tex.add(pm.k_s()),
                roughness: self.tex.add(pm.roughness()),
                bump: pm.bump_map().map(|b| self.tex.add(b) + 1).unwrap_or(0),
            },
            // Measured / Mixed / Subsurface need BSDF::sample_f or a BSSRDF, both unimplemented upstream
            // (SURVEY D8); Broken is the test-only material of Primitive::simple (D21).
            _ => return Err("material kind not supported by the B200 back end (matte, plastic)".to_string()),
        };
        self.materials.push(rec);
        Ok((self.materials.len() - 1) as u32)
    }

    // Sphere { base, radius, phi_max, z_min, z_max, theta_min, theta_max } (shape/sphere.rs:17-25),
    // Cylinder { base, radius, z_min, z_max, phi_max } (cylinder.rs:17-24),
    // Disk { base, height, radius, inner_radius, phi_max } (disk.rs:15-22).
    fn push_quadric(&mut self, s: &Shape, material: u32) {
        let base = s.base();
        let flip = (base.reverse_orientation ^ base.transform_swaps_handedness) as u32;
        let (kind, radius, z_min, z_max, phi_max, theta_min, theta_max) = match s {
            &Shape::Sphere(ref q) => { let p = q.params(); (PBRTB200_QUADRIC_SPHERE, p.radius, p.z_min, p.z_max, p.phi_max, p.theta_min, p.theta_max) }
            &Shape::Cylinder(ref q) => { let p = q.params(); (PBRTB200_QUADRIC_CYLINDER, p.radius, p.z_min, p.z_max, p.phi_max, 0.0, 0.0) }
            // disk: z_min = z_max = height, theta_min = inner_radius (include/pbrtb200.h)
            &Shape::Disk(ref q) => { let p = q.params(); (PBRTB200_QUADRIC_DISK, p.radius, p.height, p.height, p.phi_max, p.inner_radius, 0.0) }
            _ => unreachable!(),
        };
        self.spheres.push(pbrtb200_sphere80 {
            w2o: rows3(&base.world2object.m().m),
            radius, z_min, z_max, phi_max, theta_min, theta_max, material,
            flip: flip | (kind << PBRTB200_QUADRIC_KIND_SHIFT),
        });
        self.sphere_o2w.extend_from_slice(&rows3(&base.object2world.m().m));
    }

    /// Raw-pointer view for pbrtb200_upload_scene.  Valid while `self` is alive and unmodified.
    pub fn as_desc(&self) -> pbrtb200_scene {
        fn ptr<T>(v: &Vec<T>) -> *const T { if v.is_empty() { std::ptr::null() } else { v.as_ptr() } }
        pbrtb200_scene {
            nodes: ptr(&self.nodes), n_nodes: self.nodes.len() as u32,
            leaf_prim: if self.has_quadrics { ptr(&self.leaf_prim) } else { std::ptr::null() },
            n_prims: self.prim_ids.len() as u32,
            tris: ptr(&self.tris), n_tris: self.tris.len() as u32,
            spheres: ptr(&self.spheres), sphere_o2w: ptr(&self.sphere_o2w), n_spheres: self.spheres.len() as u32,
            meshes: ptr(&self.meshes), n_meshes: self.meshes.len() as u32,
            tri_uv: if self.any_uv { ptr(&self.tri_uv) } else { std::ptr::null() },
            tri_n: if self.any_n { ptr(&self.tri_n) } else { std::ptr::null() },
            tri_s: if self.any_s { ptr(&self.tri_s) } else { std::ptr::null() },
            n_attr: self.tris.len() as u32,
            materials: ptr(&self.materials), n_materials: self.materials.len() as u32,
            textures: ptr(&self.tex.textures), n_textures: self.tex.textures.len() as u32,
            lights: ptr(&self.lights), n_lights: self.lights.len() as u32,
            area_prims: ptr(&self.area_prims), n_area_prims: self.area_prims.len() as u32,
            mipmaps: ptr(&self.tex.mipmaps), n_mipmaps: self.tex.mipmaps.len() as u32,
            texels: ptr(&self.tex.texels), n_texels: (self.tex.texels.len() / 4) as u64,
        }
    }
}
