//! Lowering hooks: every `Hittable` / `Material` / `Texture` of the reference appends its POD row to a
//! `SceneBuilder`, which hands `rt1w_scene_create` the same description the C++ mirror builds
//! (`raytracing-1w_b200/host/rt1w.hpp`).  Hooks are ADDITIONS to the traits; no reference method changes.
//!
//! trait Hittable { ...; fn flatten(&self, b: &mut SceneBuilder) -> i32; }        // hittable.rs:63-72
//! trait Material { ...; fn lower(&self, b: &mut SceneBuilder) -> rt1w_material; } // material.rs:25-50
//! trait Texture  { ...; fn lower(&self, b: &mut SceneBuilder) -> i32; }           // texture.rs:8-10
use std::collections::HashMap;
use std::sync::Arc;

use crate::aabox::AABox;
use crate::aarect::{XYRect, XZRect, YZRect};
use crate::bvh::BVHNode;
use crate::camera::Camera;
use crate::constant_medium::{ConstantMedium, Isotropic};
use crate::ffi::*;
use crate::hittable::{FlipFace, Hittable, RotateY, Translate};
use crate::material::{Dielectric, DiffuseLight, Lambertian, Material, Metal};
use crate::moving_sphere::MovingSphere;
use crate::sphere::Sphere;
use crate::texture::{CheckerTexture, NoiseTexture, SolidColor, Texture};

#[derive(Default)]
pub struct SceneBuilder {
    pub nodes: Vec<rt1w_node>,
    pub children: Vec<i32>,
    pub materials: Vec<rt1w_material>,
    pub textures: Vec<rt1w_texture>,
    pub perlins: Vec<rt1w_perlin>,
    pub images: Vec<rt1w_image>,
    pub image_data: Vec<Vec<u8>>, // keeps the RGB8 buffers alive behind `images`
    pub lights: Vec<i32>,
    pub has_lights: bool,
    pub world: i32,
    material_ids: HashMap<*const (), i32>, // Arc pointer identity: a shared material lowers once
}

impl SceneBuilder {
    pub fn add_node(&mut self, type_: i32, material: i32, p: &[f64], kids: &[i32]) -> i32 {
        let mut row = rt1w_node { type_, material, child_begin: self.children.len() as i32, child_count: kids.len() as i32, p: [0.0; 10] };
        row.p[..p.len()].copy_from_slice(p);
        self.children.extend_from_slice(kids);
        self.nodes.push(row);
        self.nodes.len() as i32 - 1
    }
    pub fn material_id(&mut self, m: &Arc<Box<dyn Material>>) -> i32 {
        let key = Arc::as_ptr(m) as *const ();
        if let Some(&id) = self.material_ids.get(&key) {
            return id;
        }
        let row = m.lower(self);
        self.materials.push(row);
        let id = self.materials.len() as i32 - 1;
        self.material_ids.insert(key, id);
        id
    }
    pub fn add_texture(&mut self, t: rt1w_texture) -> i32 {
        self.textures.push(t);
        self.textures.len() as i32 - 1
    }
    /// Borrowed view for `rt1w_scene_create`; the library copies everything during the call.
    pub fn desc(&self) -> rt1w_scene_desc {
        rt1w_scene_desc {
            nodes: self.nodes.as_ptr(),
            n_nodes: self.nodes.len() as i32,
            children: self.children.as_ptr(),
            n_children: self.children.len() as i32,
            materials: self.materials.as_ptr(),
            n_materials: self.materials.len() as i32,
            textures: self.textures.as_ptr(),
            n_textures: self.textures.len() as i32,
            perlins: self.perlins.as_ptr(),
            n_perlins: self.perlins.len() as i32,
            images: self.images.as_ptr(),
            n_images: self.images.len() as i32,
            world: self.world,
            has_lights: self.has_lights as i32,
            lights: self.lights.as_ptr(),
            n_lights: self.lights.len() as i32,
        }
    }
}

fn solid(c: [f64; 3]) -> rt1w_texture {
    rt1w_texture { type_: RT1W_TEX_SOLID, odd: -1, even: -1, table: -1, color: c, scale: 0.0 }
}

// ---------------------------------------------------------------- textures (texture.rs)
impl SolidColor {
    pub fn lower_impl(&self, b: &mut SceneBuilder) -> i32 {
        b.add_texture(solid([self.color_value.0.x, self.color_value.0.y, self.color_value.0.z]))
    }
}
impl<A: Texture, B: Texture> CheckerTexture<A, B> {
    pub fn lower_impl(&self, b: &mut SceneBuilder) -> i32 {
        let (odd, even) = (self.odd.lower(b), self.even.lower(b));
        b.add_texture(rt1w_texture { type_: RT1W_TEX_CHECKER, odd, even, table: -1, color: [0.0; 3], scale: 0.0 })
    }
}
impl<const N: usize> NoiseTexture<N> {
    /// needs `pub(crate)` access to `Perlin::{ranvec, perm_x, perm_y, perm_z}` (perlin.rs:7-12); N = 256 in every scene
    pub fn lower_impl(&self, b: &mut SceneBuilder) -> i32 {
        let mut p = rt1w_perlin { ranvec: [[0.0; 3]; 256], perm_x: [0; 256], perm_y: [0; 256], perm_z: [0; 256] };
        for i in 0..256.min(N) {
            let v = self.perlin.ranvec[i];
            p.ranvec[i] = [v.x, v.y, v.z];
            p.perm_x[i] = self.perlin.perm_x[i] as i32;
            p.perm_y[i] = self.perlin.perm_y[i] as i32;
            p.perm_z[i] = self.perlin.perm_z[i] as i32;
        }
        b.perlins.push(p);
        let table = b.perlins.len() as i32 - 1;
        b.add_texture(rt1w_texture { type_: RT1W_TEX_NOISE, odd: -1, even: -1, table, color: [0.0; 3], scale: self.scale })
    }
}
/// `impl Texture for DynamicImage` (texture.rs:67-89): the decoded image as tightly packed RGB8, row 0 = top
pub fn lower_image(img: &image::DynamicImage, b: &mut SceneBuilder) -> i32 {
    let rgb = img.to_rgb8();
    let (w, h) = rgb.dimensions();
    b.image_data.push(rgb.into_raw());
    let data = b.image_data.last().unwrap();
    b.images.push(rt1w_image { rgb8: data.as_ptr(), width: w as i32, height: h as i32 });
    let table = b.images.len() as i32 - 1;
    b.add_texture(rt1w_texture { type_: RT1W_TEX_IMAGE, odd: -1, even: -1, table, color: [0.0; 3], scale: 0.0 })
}

// ---------------------------------------------------------------- materials (material.rs, constant_medium.rs:31-51)
fn material_row(type_: i32, texture: i32, albedo: [f64; 3], fuzz: f64, ir: f64) -> rt1w_material {
    rt1w_material { type_, texture, albedo, fuzz, ir }
}
impl<T: Texture> Lambertian<T> {
    pub fn lower_impl(&self, b: &mut SceneBuilder) -> rt1w_material {
        let t = self.albedo.lower(b);
        material_row(RT1W_MAT_LAMBERTIAN, t, [0.0; 3], 0.0, 0.0)
    }
}
impl Metal {
    pub fn lower_impl(&self, _b: &mut SceneBuilder) -> rt1w_material {
        material_row(RT1W_MAT_METAL, -1, [self.albedo.0.x, self.albedo.0.y, self.albedo.0.z], self.fuzz, 0.0)
    }
}
impl Dielectric {
    pub fn lower_impl(&self, _b: &mut SceneBuilder) -> rt1w_material {
        material_row(RT1W_MAT_DIELECTRIC, -1, [0.0; 3], 0.0, self.ir)
    }
}
impl<T: Texture> DiffuseLight<T> {
    pub fn lower_impl(&self, b: &mut SceneBuilder) -> rt1w_material {
        let t = self.emit.lower(b);
        material_row(RT1W_MAT_DIFFUSE_LIGHT, t, [0.0; 3], 0.0, 0.0)
    }
}
impl Isotropic {
    pub fn lower_impl(&self, b: &mut SceneBuilder) -> rt1w_material {
        let t = self.albedo.lower(b);
        material_row(RT1W_MAT_ISOTROPIC, t, [0.0; 3], 0.0, 0.0)
    }
}
/// `impl Material for ()` (material.rs:68): neither emits nor scatters
pub fn lower_null_material() -> rt1w_material {
    material_row(RT1W_MAT_NONE, -1, [0.0; 3], 0.0, 0.0)
}

// ---------------------------------------------------------------- hittables
impl Sphere {
    pub fn flatten_impl(&self, b: &mut SceneBuilder) -> i32 {
        let m = b.material_id(&self.material);
        b.add_node(RT1W_NODE_SPHERE, m, &[self.center.x, self.center.y, self.center.z, self.radius], &[])
    }
}
impl MovingSphere {
    pub fn flatten_impl(&self, b: &mut SceneBuilder) -> i32 {
        let m = b.material_id(&self.material);
        let (c0, c1) = (self.center0, self.center1);
        b.add_node(RT1W_NODE_MOVING_SPHERE, m, &[c0.x, c0.y, c0.z, c1.x, c1.y, c1.z, self.time0, self.time1, self.radius], &[])
    }
}
impl XYRect {
    pub fn flatten_impl(&self, b: &mut SceneBuilder) -> i32 {
        let m = b.material_id(&self.material);
        b.add_node(RT1W_NODE_XY_RECT, m, &[self.x0, self.x1, self.y0, self.y1, self.k], &[])
    }
}
impl XZRect {
    pub fn flatten_impl(&self, b: &mut SceneBuilder) -> i32 {
        let m = b.material_id(&self.material);
        b.add_node(RT1W_NODE_XZ_RECT, m, &[self.x0, self.x1, self.z0, self.z1, self.k], &[])
    }
}
impl YZRect {
    pub fn flatten_impl(&self, b: &mut SceneBuilder) -> i32 {
        let m = b.material_id(&self.material);
        b.add_node(RT1W_NODE_YZ_RECT, m, &[self.y0, self.y1, self.z0, self.z1, self.k], &[])
    }
}
impl AABox {
    /// needs the box's material kept next to `box_min/box_max` (aabox.rs:16-20 stores only the six sides)
    pub fn flatten_impl(&self, b: &mut SceneBuilder) -> i32 {
        let m = b.material_id(&self.material);
        let (p0, p1) = (self.box_min, self.box_max);
        b.add_node(RT1W_NODE_AABOX, m, &[p0.x, p0.y, p0.z, p1.x, p1.y, p1.z], &[])
    }
}
impl<T: Hittable> Translate<T> {
    pub fn flatten_impl(&self, b: &mut SceneBuilder) -> i32 {
        let c = self.hittable.flatten(b);
        b.add_node(RT1W_NODE_TRANSLATE, -1, &[self.offset.x, self.offset.y, self.offset.z], &[c])
    }
}
impl<T: Hittable> RotateY<T> {
    /// needs `angle: Deg<Float>`, `time0`, `time1` kept by `RotateY::new` (hittable.rs:158-203 keeps only sin/cos and the box)
    pub fn flatten_impl(&self, b: &mut SceneBuilder) -> i32 {
        let c = self.hittable.flatten(b);
        b.add_node(RT1W_NODE_ROTATE_Y, -1, &[self.angle.0, self.time0, self.time1], &[c])
    }
}
impl<T: Hittable> FlipFace<T> {
    pub fn flatten_impl(&self, b: &mut SceneBuilder) -> i32 {
        let c = self.0.flatten(b);
        b.add_node(RT1W_NODE_FLIP_FACE, -1, &[], &[c])
    }
}
impl<T: Hittable> ConstantMedium<T> {
    /// density = -1 / neg_inv_density (constant_medium.rs:26); the phase function is always an Isotropic
    pub fn flatten_impl(&self, b: &mut SceneBuilder) -> i32 {
        let c = self.boundary.flatten(b);
        let m = b.material_id(&self.phase_function);
        b.add_node(RT1W_NODE_CONSTANT_MEDIUM, m, &[-1.0 / self.neg_inv_density], &[c])
    }
}
impl BVHNode {
    /// needs `child_ids: Vec<i32>` collected by `BVHNode::new` BEFORE it sorts and splits `objects`
    /// (bvh.rs:54-103), plus its `time0/time1`.  The random median split is discarded: the device builds
    /// one SAH (or LBVH) tree over all leaves of the scene.
    pub fn flatten_impl(&self, b: &mut SceneBuilder) -> i32 {
        let kids: Vec<i32> = self.flat_children.iter().map(|h| h.flatten(b)).collect();
        b.add_node(RT1W_NODE_BVH, -1, &[self.time0, self.time1], &kids)
    }
}

// ---------------------------------------------------------------- camera (camera.rs:8-19)
impl Camera {
    pub fn as_ffi(&self) -> rt1w_camera {
        let p = |v: cgmath::Point3<f64>| [v.x, v.y, v.z];
        let v3 = |v: cgmath::Vector3<f64>| [v.x, v.y, v.z];
        rt1w_camera {
            origin: p(self.origin),
            lower_left_corner: p(self.lower_left_corner),
            horizontal: v3(self.horizontal),
            vertical: v3(self.vertical),
            u: v3(self.u),
            v: v3(self.v),
            w: v3(self.w),
            lens_radius: self.lens_radius,
            time0: self.time0,
            time1: self.time1,
        }
    }
}
