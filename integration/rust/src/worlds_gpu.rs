//! `worlds.rs` written against `gpu::SceneSink`: every recipe makes the SAME `rng` calls in the SAME order as
//! `World::build` (src/worlds.rs), and records one sink entry where the reference calls a constructor.  The
//! descriptions must hash (`SceneSink::hash`) to tests/golden/scene_hashes.json of the B200 repository — the values its
//! C++ builder (csrc/worlds.cpp) and its oracle produce for the same seed.
//!
//! NOT compiled in this repository (no Rust toolchain in the image).
use crate::gpu::{SceneSink, RT_BG_BLACK, RT_BG_GRADIENT};
use crate::vec::{Color, Point3};
use rand::Rng;

fn rnd01(rng: &mut dyn rand::RngCore) -> f64 {
    rng.gen_range(0.0..1.0)
}

/// (root node, background kind) of world `name`; `earthmap` is `image::open("earthmap.jpg").unwrap().to_rgb8()` for the
/// two worlds that use it (worlds.rs:180, :441).
pub fn describe(name: &str, rng: &mut dyn rand::RngCore, earthmap: Option<&image::RgbImage>, s: &mut SceneSink) -> (i32, i32) {
    match name {
        "simple" => (simple(rng, s), RT_BG_GRADIENT),
        "random" => (random(rng, s, false), RT_BG_GRADIENT),
        "random_chk" => (random(rng, s, true), RT_BG_GRADIENT),
        "two_spheres" => (two_spheres(rng, s, false), RT_BG_GRADIENT),
        "simple_light" => (two_spheres(rng, s, true), RT_BG_BLACK),
        "cornell_box" => (cornell(s, false), RT_BG_BLACK),
        "cornell_smoke" => (cornell(s, true), RT_BG_BLACK),
        "earth" => (earth(earthmap.expect("earthmap.jpg"), s), RT_BG_GRADIENT),
        "debug_perlin" => (debug_perlin(rng, s), RT_BG_GRADIENT),
        "final_scene" => (final_scene(rng, earthmap.expect("earthmap.jpg"), s), RT_BG_BLACK),
        _ => panic!("unknown world {}", name),
    }
}

fn simple(rng: &mut dyn rand::RngCore, s: &mut SceneSink) -> i32 {
    // worlds.rs:41-59
    let t = s.solid(0.8, 0.8, 0.0);
    let ground = s.lambertian(t);
    let t = s.solid(0.1, 0.3, 0.5);
    let center = s.lambertian(t);
    let left = s.dielectric(1.5);
    let right = s.metal([0.8, 0.6, 0.2], 0.0);
    let items = [
        s.sphere([0.0, -100.5, -1.0], 100.0, ground),
        s.sphere([0.0, 0.0, -1.0], 0.5, center),
        s.sphere([-1.0, 0.0, -1.0], 0.5, left),
        s.sphere([-1.0, 0.0, -1.0], -0.4, left),
        s.sphere([1.0, 0.0, -1.0], 0.5, right),
    ];
    s.bvh(&items, rng)
}

fn random(rng: &mut dyn rand::RngCore, s: &mut SceneSink, checker_ground: bool) -> i32 {
    // worlds.rs:79-112 (random), :128-162 (random_chk)
    let mut items = Vec::new();
    let ground_tex = if checker_ground {
        let odd = s.solid(0.2, 0.3, 0.1);
        let even = s.solid(0.9, 0.9, 0.9);
        s.checker(odd, even)
    } else {
        s.solid(0.5, 0.5, 0.5)
    };
    let ground = s.lambertian(ground_tex);
    items.push(s.sphere([0.0, -1000.0, 0.0], 1000.0, ground));
    for a in -11..11 {
        for b in -11..11 {
            let choose_mat = rnd01(rng);
            let center = Point3::new(a as f64 + 0.9 * rnd01(rng), 0.2, b as f64 + 0.9 * rnd01(rng));
            if (center - Point3::new(4.0, 0.2, 0.0)).length() > 0.9 {
                let mat = if choose_mat < 0.8 {
                    let albedo = Color::random_unit(rng) * Color::random_unit(rng);
                    let t = s.solid(albedo.e[0], albedo.e[1], albedo.e[2]);
                    s.lambertian(t)
                } else if choose_mat < 0.95 {
                    let albedo = Color::random(0.5, 1.0, rng);
                    let fuzz = rng.gen_range(0.0..0.5);
                    s.metal(albedo.e, fuzz)
                } else {
                    s.dielectric(1.5)
                };
                items.push(s.sphere(center.e, 0.2, mat));
            }
        }
    }
    let glass = s.dielectric(1.5);
    items.push(s.sphere([0.0, 1.0, 0.0], 1.0, glass));
    let t = s.solid(0.4, 0.2, 0.1);
    let brown = s.lambertian(t);
    items.push(s.sphere([-4.0, 1.0, 0.0], 1.0, brown));
    let steel = s.metal([0.7, 0.6, 0.5], 0.0);
    items.push(s.sphere([4.0, 1.0, 0.0], 1.0, steel));
    s.bvh(&items, rng)
}

fn earth(img: &image::RgbImage, s: &mut SceneSink) -> i32 {
    // worlds.rs:179-186
    let t = s.image(img);
    let surface = s.lambertian(t);
    s.sphere([0.0, 0.0, 0.0], 2.0, surface)
}

fn two_spheres(rng: &mut dyn rand::RngCore, s: &mut SceneSink, lights: bool) -> i32 {
    // worlds.rs:203-210 (two_spheres), :227-239 (simple_light)
    let pertext = s.noise(4.0, rng);
    let mat = s.lambertian(pertext);
    let mut items = vec![s.sphere([0.0, -1000.0, 0.0], 1000.0, mat)];
    let mat = s.lambertian(pertext); // Lambertian::new(pertext): a second material over the same (cloned) texture
    items.push(s.sphere([0.0, 2.0, 0.0], 2.0, mat));
    if lights {
        let t = s.solid(0.0, 7.0, 0.0);
        let green = s.diffuse_light(t);
        items.push(s.xy_rect(3.0, 5.0, 1.0, 3.0, -2.0, green));
        let t = s.solid(7.0, 0.0, 0.0);
        let red = s.diffuse_light(t);
        items.push(s.sphere([0.0, 6.0, 0.0], 1.5, red));
    }
    s.list(&items)
}

fn cornell(s: &mut SceneSink, smoke: bool) -> i32 {
    // worlds.rs:259-287 (cornell_box), :308-335 (cornell_smoke)
    let t = s.solid(0.65, 0.05, 0.05);
    let red = s.lambertian(t);
    let t = s.solid(0.73, 0.73, 0.73);
    let white = s.lambertian(t);
    let t = s.solid(0.12, 0.45, 0.15);
    let green = s.lambertian(t);
    let t = s.solid(7.0, 7.0, 7.0);
    let light = s.diffuse_light(t);
    let mut items = vec![
        s.yz_rect(0.0, 555.0, 0.0, 555.0, 555.0, green),
        s.yz_rect(0.0, 555.0, 0.0, 555.0, 0.0, red),
        s.xz_rect(113.0, 443.0, 127.0, 432.0, 554.0, light),
        s.xz_rect(0.0, 555.0, 0.0, 555.0, 0.0, white),
        s.xz_rect(0.0, 555.0, 0.0, 555.0, 555.0, white),
        s.xy_rect(0.0, 555.0, 0.0, 555.0, 555.0, white),
    ];
    let b = s.block([0.0, 0.0, 0.0], [165.0, 330.0, 165.0], white);
    let b = s.rotate(1, 15.0, b);
    let large = s.translate([265.0, 0.0, 295.0], b);
    items.push(if smoke { s.medium(large, 0.01, [0.0, 0.0, 0.0]) } else { large });
    let b = s.block([0.0, 0.0, 0.0], [165.0, 165.0, 165.0], white);
    let b = s.rotate(1, -18.0, b);
    let small = s.translate([130.0, 0.0, 65.0], b);
    items.push(if smoke { s.medium(small, 0.01, [1.0, 1.0, 1.0]) } else { small });
    s.list(&items)
}

fn debug_perlin(rng: &mut dyn rand::RngCore, s: &mut SceneSink) -> i32 {
    // worlds.rs:355-365
    let pertext = s.noise(0.1, rng);
    let mat = s.lambertian(pertext);
    let ball = s.sphere([278.0, 278.0, 0.0], 80.0, mat);
    s.list(&[ball])
}

fn final_scene(rng: &mut dyn rand::RngCore, earthmap: &image::RgbImage, s: &mut SceneSink) -> i32 {
    // worlds.rs:386-468
    let mut items = Vec::new();
    {
        let t = s.solid(9.0, 9.0, 9.0);
        let light = s.diffuse_light(t);
        items.push(s.xz_rect(123.0, 423.0, 147.0, 412.0, 554.0, light));
    }
    {
        let t = s.solid(0.48, 0.83, 0.53);
        let ground = s.lambertian(t);
        let mut blocks = Vec::new();
        for i in 0..20 {
            for j in 0..20 {
                let w = 100.0;
                let x0 = -1000.0 + (i as f64) * w;
                let z0 = -1000.0 + (j as f64) * w;
                let y1 = rng.gen_range(1.0..70.0);
                blocks.push(s.block([x0, 0.0, z0], [x0 + w, y1, z0 + w], ground));
            }
        }
        items.push(s.bvh(&blocks, rng));
    }
    let t = s.solid(0.7, 0.3, 0.1);
    let gold = s.lambertian(t);
    items.push(s.sphere([400.0, 400.0, 400.0], 50.0, gold));
    let glass = s.dielectric(1.5);
    items.push(s.sphere([260.0, 150.0, 45.0], 50.0, glass));
    let metal = s.metal([0.8, 0.8, 0.9], 1.0);
    items.push(s.sphere([0.0, 150.0, 145.0], 50.0, metal));
    {
        let glass = s.dielectric(1.5);
        let boundary = s.sphere([360.0, 150.0, 145.0], 70.0, glass);
        items.push(boundary); // boundary.clone(): the glass shell itself ...
        items.push(s.medium(boundary, 0.2, [0.2, 0.4, 0.9])); // ... and the smoke inside it
    }
    {
        let glass = s.dielectric(1.5);
        let boundary = s.sphere([0.0, 0.0, 0.0], 1000.0, glass);
        items.push(s.medium(boundary, 0.0001, [1.0, 1.0, 1.0]));
    }
    {
        let t = s.image(earthmap);
        let surface = s.lambertian(t);
        items.push(s.sphere([400.0, 200.0, 400.0], 100.0, surface));
    }
    {
        let pertext = s.noise(0.1, rng);
        let mat = s.lambertian(pertext);
        items.push(s.sphere([220.0, 280.0, 300.0], 80.0, mat));
    }
    {
        let t = s.solid(0.73, 0.73, 0.73);
        let white = s.lambertian(t);
        let mut foam = Vec::new();
        for _ in 0..1000 {
            let c = Point3::random(0.0, 165.0, rng);
            foam.push(s.sphere(c.e, 10.0, white));
        }
        let foam = s.bvh(&foam, rng);
        let foam = s.rotate(1, 15.0, foam);
        items.push(s.translate([-100.0, 270.0, 395.0], foam));
    }
    s.list(&items)
}
