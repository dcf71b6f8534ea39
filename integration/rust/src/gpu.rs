//! FFI over include/rt_b200.h: the GPU replacement of `Renderer::new_with_rng(..).render(..)` (src/raytrace.rs:151-186).
//! NOT compiled in this repository (no Rust toolchain in the image); kept in step with the header by hand.
#![allow(non_camel_case_types, dead_code)]
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

pub const RT_NODE_SPHERE: i32 = 1;
pub const RT_NODE_XYRECT: i32 = 2;
pub const RT_NODE_XZRECT: i32 = 3;
pub const RT_NODE_YZRECT: i32 = 4;
pub const RT_NODE_BLOCK: i32 = 5;
pub const RT_NODE_TRANSLATE: i32 = 6;
pub const RT_NODE_ROTATE: i32 = 7;
pub const RT_NODE_MEDIUM: i32 = 8;
pub const RT_NODE_BVH: i32 = 9;
pub const RT_NODE_LIST: i32 = 10;
pub const RT_MAT_LAMBERTIAN: i32 = 1;
pub const RT_MAT_METAL: i32 = 2;
pub const RT_MAT_DIELECTRIC: i32 = 3;
pub const RT_MAT_DIFFUSE_LIGHT: i32 = 4;
pub const RT_MAT_ISOTROPIC: i32 = 5;
pub const RT_TEX_SOLID: i32 = 1;
pub const RT_TEX_CHECKER: i32 = 2;
pub const RT_TEX_NOISE: i32 = 3;
pub const RT_TEX_IMAGE: i32 = 4;

#[repr(C)] #[derive(Clone, Copy)]
pub struct RtNode { pub kind: i32, pub material: i32, pub first_child: i32, pub child_count: i32, pub axis: i32, pub reserved: i32, pub f: [f64; 8] }
#[repr(C)] #[derive(Clone, Copy)]
pub struct RtMaterial { pub kind: i32, pub texture: i32, pub albedo: [f64; 3], pub fuzz: f64, pub ior: f64 }
#[repr(C)] #[derive(Clone, Copy)]
pub struct RtTexture { pub kind: i32, pub a: i32, pub b: i32, pub reserved: i32, pub color: [f64; 3], pub scale: f64 }
#[repr(C)]
pub struct RtPerlin { pub ranvec: [[f64; 3]; 1024], pub perm_x: [i32; 1024], pub perm_y: [i32; 1024], pub perm_z: [i32; 1024] }
#[repr(C)]
pub struct RtImage { pub width: i32, pub height: i32, pub rgb: *const u8 }
#[repr(C)]
pub struct RtSceneDesc {
    pub root: i32, pub background_kind: i32, pub background_top: [f64; 3], pub background_bottom: [f64; 3],
    pub n_nodes: i32, pub n_children: i32, pub n_materials: i32, pub n_textures: i32, pub n_perlins: i32, pub n_images: i32,
    pub nodes: *const RtNode, pub children: *const i32, pub materials: *const RtMaterial, pub textures: *const RtTexture,
    pub perlins: *const RtPerlin, pub images: *const RtImage,
}
#[repr(C)]
pub struct RtCamera { pub lookfrom: [f64; 3], pub lookat: [f64; 3], pub vup: [f64; 3], pub vfov_deg: f64, pub aspect_ratio: f64, pub aperture: f64, pub focus_dist: f64 }
#[repr(C)]
pub struct RtParams { pub width: i32, pub height: i32, pub samples_per_pixel: i32, pub max_depth: i32, pub seed: u64, pub sample_begin: i32,
                      pub sample_count: i32, pub pipeline: i32, pub device: i32, pub samples_per_item: i32, pub reserved: i32 }
#[repr(C)] #[derive(Default)]
pub struct RtStats { pub paths: u64, pub rays: u64, pub device_ms: f64, pub kernel_launches: i32, pub pipeline_used: i32 }
pub enum RtScene {}
pub type RtProgressFn = Option<extern "C" fn(done: c_int, total: c_int, user: *mut c_void)>;

extern "C" {
    pub fn rt_last_error() -> *const c_char;
    pub fn rt_device_count() -> c_int;
    pub fn rt_scene_hash(desc: *const RtSceneDesc, out: *mut u8) -> c_int;
    pub fn rt_scene_create(desc: *const RtSceneDesc, device: c_int, out: *mut *mut RtScene) -> c_int;
    pub fn rt_scene_destroy(scene: *mut RtScene);
    pub fn rt_render(scene: *const RtScene, cam: *const RtCamera, params: *const RtParams, accum_rgb: *mut f32, rgb: *mut i32,
                     cb: RtProgressFn, user: *mut c_void, stats: *mut RtStats) -> c_int;
    pub fn rt_render_multi(scenes: *const *mut RtScene, n_scenes: i32, cam: *const RtCamera, params: *const RtParams, accum_rgb: *mut f32,
                           rgb: *mut i32, cb: RtProgressFn, user: *mut c_void, stats: *mut RtStats) -> c_int;
}

fn check(code: c_int) {
    if code != 0 {
        // the reference panics on misuse (unwrap); so does the shim, with the library's message
        panic!("rt_b200: {}", unsafe { CStr::from_ptr(rt_last_error()) }.to_string_lossy());
    }
}

/// What `World::build` fills while it constructs the `Box<dyn Hittable>` tree: one entry per constructor call, in
/// call order, values exactly as the Rust code holds them (f64).  `Arc::clone` shares the entry.
#[derive(Default)]
pub struct SceneSink {
    pub nodes: Vec<RtNode>, pub children: Vec<i32>, pub materials: Vec<RtMaterial>, pub textures: Vec<RtTexture>,
    pub perlins: Vec<Box<RtPerlin>>, pub images: Vec<(i32, i32, Vec<u8>)>,
}
impl SceneSink {
    fn node(&mut self, kind: i32, material: i32, first_child: i32, child_count: i32, axis: i32, f: &[f64]) -> i32 {
        let mut a = [0.0; 8];
        a[..f.len()].copy_from_slice(f);
        self.nodes.push(RtNode { kind, material, first_child, child_count, axis, reserved: 0, f: a });
        self.nodes.len() as i32 - 1
    }
    pub fn solid(&mut self, c: [f64; 3]) -> i32 { self.textures.push(RtTexture { kind: RT_TEX_SOLID, a: -1, b: -1, reserved: 0, color: c, scale: 0.0 }); self.textures.len() as i32 - 1 }
    pub fn lambertian(&mut self, tex: i32) -> i32 { self.materials.push(RtMaterial { kind: RT_MAT_LAMBERTIAN, texture: tex, albedo: [0.0; 3], fuzz: 0.0, ior: 0.0 }); self.materials.len() as i32 - 1 }
    pub fn metal(&mut self, albedo: [f64; 3], fuzz: f64) -> i32 { self.materials.push(RtMaterial { kind: RT_MAT_METAL, texture: -1, albedo, fuzz, ior: 0.0 }); self.materials.len() as i32 - 1 }
    pub fn dielectric(&mut self, ior: f64) -> i32 { self.materials.push(RtMaterial { kind: RT_MAT_DIELECTRIC, texture: -1, albedo: [0.0; 3], fuzz: 0.0, ior }); self.materials.len() as i32 - 1 }
    pub fn sphere(&mut self, c: [f64; 3], r: f64, mat: i32) -> i32 { self.node(RT_NODE_SPHERE, mat, -1, 0, 0, &[c[0], c[1], c[2], r]) }
    pub fn xz_rect(&mut self, x0: f64, x1: f64, z0: f64, z1: f64, y: f64, mat: i32) -> i32 { self.node(RT_NODE_XZRECT, mat, -1, 0, 0, &[x0, x1, z0, z1, y]) }
    pub fn block(&mut self, p0: [f64; 3], p1: [f64; 3], mat: i32) -> i32 { self.node(RT_NODE_BLOCK, mat, -1, 0, 0, &[p0[0], p0[1], p0[2], p1[0], p1[1], p1[2]]) }
    pub fn translate(&mut self, off: [f64; 3], child: i32) -> i32 { self.node(RT_NODE_TRANSLATE, -1, child, 0, 0, &off) }
    pub fn rotate(&mut self, axis: i32, degrees: f64, child: i32) -> i32 { self.node(RT_NODE_ROTATE, -1, child, 0, axis, &[degrees]) }
    pub fn medium(&mut self, boundary: i32, density: f64, isotropic: i32) -> i32 { self.node(RT_NODE_MEDIUM, isotropic, boundary, 0, 0, &[density]) }
    /// `BHV::new(items, rng)` still draws its split axes from `rng` (bhv.rs:127) so that whatever is built next sees the
    /// same stream; only the member list is recorded (the device builds its own SAH BVH).
    pub fn group(&mut self, kind: i32, items: &[i32]) -> i32 {
        let first = self.children.len() as i32;
        self.children.extend_from_slice(items);
        self.node(kind, -1, first, items.len() as i32, 0, &[])
    }
    // xy/yz rects, checker / noise / image textures, diffuse light, isotropic: same pattern (see csrc/worlds.cpp)
}

/// The GPU arm of do_tracing: same inputs as `Renderer::new_with_rng`, same output as `Renderer::render`
/// (H rows of W (r, g, b), row 0 at the BOTTOM).
pub fn render_gpu(sink: &SceneSink, root: i32, background_kind: i32, top: [f64; 3], bottom: [f64; 3], cam: &RtCamera,
                  width: usize, height: usize, spp: i32, max_depth: i32, seed: u64, gpus: usize) -> Vec<Vec<(i32, i32, i32)>> {
    let perlins: Vec<*const RtPerlin> = sink.perlins.iter().map(|p| &**p as *const RtPerlin).collect();
    let _ = perlins; // (contiguous RtPerlin array elided: copy the boxes into one Vec<RtPerlin> before the call)
    let images: Vec<RtImage> = sink.images.iter().map(|(w, h, px)| RtImage { width: *w, height: *h, rgb: px.as_ptr() }).collect();
    let desc = RtSceneDesc {
        root, background_kind, background_top: top, background_bottom: bottom,
        n_nodes: sink.nodes.len() as i32, n_children: sink.children.len() as i32, n_materials: sink.materials.len() as i32,
        n_textures: sink.textures.len() as i32, n_perlins: 0, n_images: images.len() as i32,
        nodes: sink.nodes.as_ptr(), children: sink.children.as_ptr(), materials: sink.materials.as_ptr(), textures: sink.textures.as_ptr(),
        perlins: std::ptr::null(), images: images.as_ptr(),
    };
    let mut scenes: Vec<*mut RtScene> = vec![std::ptr::null_mut(); gpus];
    for (g, s) in scenes.iter_mut().enumerate() { check(unsafe { rt_scene_create(&desc, g as c_int, s) }); }
    let prm = RtParams { width: width as i32, height: height as i32, samples_per_pixel: spp, max_depth, seed, sample_begin: 0, sample_count: 0,
                         pipeline: 0, device: -1, samples_per_item: 0, reserved: 0 };
    let mut rgb = vec![0i32; 3 * width * height];
    let mut stats = RtStats::default();
    check(unsafe { rt_render_multi(scenes.as_ptr(), gpus as i32, cam, &prm, std::ptr::null_mut(), rgb.as_mut_ptr(), None, std::ptr::null_mut(), &mut stats) });
    for s in scenes { unsafe { rt_scene_destroy(s) }; }
    (0..height).map(|j| (0..width).map(|i| { let k = 3 * (j * width + i); (rgb[k], rgb[k + 1], rgb[k + 2]) }).collect()).collect()
}
