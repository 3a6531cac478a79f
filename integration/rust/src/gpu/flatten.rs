// Sketch of the flatten shim: not compiled in the build image (no Rust toolchain).
// It lives inside the crate because the fields it reads are private.
use super::ffi::*;

// short names used in the sketch
pub type Node32 = pbrtb200_node32;
pub type Tri48 = pbrtb200_tri48;
pub type Sphere80 = pbrtb200_sphere80;
pub type MeshRec = pbrtb200_mesh;
pub type TextureRec = pbrtb200_texture;
pub type MaterialRec = pbrtb200_material;
pub type LightRec = pbrtb200_light;
pub type SceneDesc = pbrtb200_scene;

pub struct FlatScene { nodes: Vec<Node32>, leaf_prim: Vec<u32>, tris: Vec<Tri48>,
                       spheres: Vec<Sphere80>, sphere_o2w: Vec<f32>, meshes: Vec<MeshRec>, /* … */ }

impl FlatScene {
    pub fn from_scene(scene: &Scene) -> FlatScene {
        let bvh: &BVHAccelerator = scene.aggregate_bvh();          // Aggregate::BVH(..) arm
        // 1. nodes: a straight copy of BVHAccelerator.nodes (bvh.rs:260-272, depth-first order)
        let nodes = bvh.nodes.iter().map(|n| match n {
            PackedBVHNode::Leaf  { bounds, prim_offset, num_prims } =>
                Node32 { bmin: bounds.p_min.into(), bmax: bounds.p_max.into(),
                         offset: *prim_offset as u32, count: *num_prims as u16, axis: 0, is_leaf: 1 },
            PackedBVHNode::Inner { bounds, second_child_offset, axis } =>
                Node32 { bmin: bounds.p_min.into(), bmax: bounds.p_max.into(),
                         offset: *second_child_offset as u32, count: 0, axis: *axis as u8, is_leaf: 0 },
        }).collect();
        // 2. ordered primitives (bvh.rs:331): triangle -> Tri48 with world-space vertices in
        //    Triangle.v order (already refine-reversed, mesh.rs:329-331); sphere -> Sphere80 with
        //    base.world2object rows 0..2 and base.object2world rows 0..2.
        //    flip = reverse_orientation ^ transform_swaps_handedness (shape/mod.rs:33-54).
        // 3. materials/textures/lights -> the tagged records; PointLight.light_pos, SpotLight's
        //    world_to_light + cos_total_width/cos_falloff_start (light/spot.rs:24-35).
        //    Textures are `Arc<dyn Texture<T>>` trait objects (texture/mod.rs:36-46), which cannot be
        //    matched on from outside: the shim adds one method to the crate's own trait,
        //        fn flatten(&self, table: &mut TexTable) -> i32   // pushes its TextureRec, returns its index
        //    implemented per texture type (Constant -> kind 0 + value; Checkerboard -> kind 1, the
        //    mapping's `flatten_mapping()` into map_kind/map, children flattened first; Scale / Mix /
        //    Bilerp / Dots / FBm / Wrinkled / ImageTexture likewise), deduplicated by Arc pointer, and
        //    the same for `TextureMapping2D` / `TextureMapping3D` (mapping2d.rs:36-47, mapping3d.rs:30-41).
        //    MatteMaterial { k_d, sigma, bump_map } / PlasticMaterial { k_d, k_s, roughness, bump_map }
        //    -> MaterialRec { kind, kd, sigma | ks, roughness, bump = index + 1 or 0 }.
        // 4. `user` of each Tri48 / the side table keeps Primitive::get_id() so results can be
        //    reported in the crate's own ids.
        /* … */
    }
    pub fn as_desc(&self) -> SceneDesc { /* raw pointers into the Vecs */ }
}
