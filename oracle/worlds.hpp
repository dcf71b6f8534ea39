// ORACLE — test infrastructure only.  Nothing under oracle/ is part of the product path.
//
// The reference's ten scene recipes (src/worlds.rs) restated against the oracle's object tree.  Every
// recipe builds TWO things side by side, in the reference's construction order:
//   * the Hittable tree the oracle renderer traces (what World::build returns), and
//   * the scene description (RtSceneDesc, include/rt_b200.h) whose canonical hash is the
//     "bit-exact world" check against the product's own builder.
// Emission rule shared with the product builder: a desc entry is appended at the point where the
// reference calls the constructor (`Sphere::new`, `Lambertian::new`, ...); `.clone()` shares the entry.
#pragma once
#include <cstring>
#include <string>

#include "../include/rt_b200.h"
#include "scene.hpp"

namespace orc {

struct WorldOut {
    std::shared_ptr<Hittable> root;
    int background_kind = RT_BG_BLACK;
    Vec3 lookfrom, lookat;
    double vfov = 20.0;
    uint64_t draws = 0;  // next_u64 calls consumed by build
    // owning storage of the description
    std::vector<RtNode> nodes;
    std::vector<int32_t> children;
    std::vector<RtMaterial> materials;
    std::vector<RtTexture> textures;
    std::vector<RtPerlin> perlins;
    std::vector<RtImage> images;
    std::vector<std::shared_ptr<std::vector<uint8_t>>> image_data;
    std::vector<std::shared_ptr<Perlin>> perlin_objs;
    std::vector<std::vector<int>> bvh_axes;  // instrumentation: axis draws of every BVH built, in build order
    std::vector<HittablePtr> by_node;        // description node index -> tree object
    int32_t root_node = -1;
    RtSceneDesc desc;  // points into the vectors above; valid after finish()

    void finish() {
        std::memset(&desc, 0, sizeof(desc));
        desc.root = root_node;
        desc.background_kind = background_kind;
        // GradientBackground::default(): top = (0.5,0.7,1.0), bottom = white (raytrace.rs:21-26)
        double top[3] = {0.5, 0.7, 1.0}, bot[3] = {1.0, 1.0, 1.0};
        for (int i = 0; i < 3; i++) {
            desc.background_top[i] = background_kind == RT_BG_GRADIENT ? top[i] : 0.0;
            desc.background_bottom[i] = background_kind == RT_BG_GRADIENT ? bot[i] : 0.0;
        }
        desc.n_nodes = (int32_t)nodes.size();
        desc.n_children = (int32_t)children.size();
        desc.n_materials = (int32_t)materials.size();
        desc.n_textures = (int32_t)textures.size();
        desc.n_perlins = (int32_t)perlins.size();
        desc.n_images = (int32_t)images.size();
        desc.nodes = nodes.data();
        desc.children = children.data();
        desc.materials = materials.data();
        desc.textures = textures.data();
        desc.perlins = perlins.data();
        desc.images = images.data();
    }
};

// a handle that carries the tree object and its description index together
struct Obj {
    HittablePtr h;
    int32_t node;
};
struct Mat {
    MaterialPtr m;
    int32_t index;
};
struct Tex {
    std::shared_ptr<Texture> t;
    int32_t index;
};

struct Recipe {
    WorldOut& w;
    Pcg64& rng;
    Recipe(WorldOut& out, Pcg64& r) : w(out), rng(r) {}

    // ---- textures ----
    Tex solid(double r, double g, double b) {  // SolidColor::new / from_color
        RtTexture t;
        std::memset(&t, 0, sizeof(t));
        t.kind = RT_TEX_SOLID;
        t.a = t.b = -1;
        t.color[0] = r; t.color[1] = g; t.color[2] = b;
        w.textures.push_back(t);
        return Tex{std::make_shared<SolidColor>(Color(r, g, b)), (int32_t)w.textures.size() - 1};
    }
    Tex solid(const Color& c) { return solid(c.x(), c.y(), c.z()); }
    Tex checker(const Tex& odd, const Tex& even) {  // Checker::new(odd, even)
        RtTexture t;
        std::memset(&t, 0, sizeof(t));
        t.kind = RT_TEX_CHECKER;
        t.a = odd.index; t.b = even.index;
        w.textures.push_back(t);
        return Tex{std::make_shared<Checker>(odd.t, even.t), (int32_t)w.textures.size() - 1};
    }
    Tex noise(double scale) {  // NoiseTexture::new(scale, rng) -> Perlin::new(rng)
        auto p = std::make_shared<Perlin>(rng);
        w.perlin_objs.push_back(p);
        RtPerlin rp;
        for (int i = 0; i < Perlin::N; i++) {
            for (int c = 0; c < 3; c++) rp.ranvec[i][c] = p->ranvec[i].e[c];
            rp.perm_x[i] = (int32_t)p->perm_x[i];
            rp.perm_y[i] = (int32_t)p->perm_y[i];
            rp.perm_z[i] = (int32_t)p->perm_z[i];
        }
        w.perlins.push_back(rp);
        RtTexture t;
        std::memset(&t, 0, sizeof(t));
        t.kind = RT_TEX_NOISE;
        t.a = (int32_t)w.perlins.size() - 1; t.b = -1;
        t.scale = scale;
        w.textures.push_back(t);
        return Tex{std::make_shared<NoiseTexture>(p, scale), (int32_t)w.textures.size() - 1};
    }
    Tex image(const uint8_t* rgb, int width, int height) {  // image_texture::Image::new(img.to_rgb8())
        auto data = std::make_shared<std::vector<uint8_t>>(rgb, rgb + (size_t)3 * width * height);
        w.image_data.push_back(data);
        RtImage im;
        im.width = width; im.height = height; im.rgb = data->data();
        w.images.push_back(im);
        RtTexture t;
        std::memset(&t, 0, sizeof(t));
        t.kind = RT_TEX_IMAGE;
        t.a = (int32_t)w.images.size() - 1; t.b = -1;
        w.textures.push_back(t);
        return Tex{std::make_shared<ImageTexture>(width, height, data), (int32_t)w.textures.size() - 1};
    }

    // ---- materials ----
    Mat push_mat(MaterialPtr m, int kind, int32_t tex, const Color& albedo, double fuzz, double ior) {
        RtMaterial d;
        std::memset(&d, 0, sizeof(d));
        d.kind = kind; d.texture = tex;
        for (int i = 0; i < 3; i++) d.albedo[i] = albedo.e[i];
        d.fuzz = fuzz; d.ior = ior;
        w.materials.push_back(d);
        m->desc_index = (int32_t)w.materials.size() - 1;
        return Mat{m, m->desc_index};
    }
    Mat lambertian(const Tex& t) { return push_mat(std::make_shared<Lambertian>(t.t), RT_MAT_LAMBERTIAN, t.index, Color(), 0, 0); }
    Mat metal(const Color& a, double fuzz) { return push_mat(std::make_shared<Metal>(a, fuzz), RT_MAT_METAL, -1, a, fuzz, 0); }
    Mat dielectric(double ior) { return push_mat(std::make_shared<Dielectric>(ior), RT_MAT_DIELECTRIC, -1, Color(), 0, ior); }
    Mat diffuse_light(const Tex& t) { return push_mat(std::make_shared<DiffuseLight>(t.t), RT_MAT_DIFFUSE_LIGHT, t.index, Color(), 0, 0); }
    Mat isotropic(const Tex& t) { return push_mat(std::make_shared<Isotropic>(t.t), RT_MAT_ISOTROPIC, t.index, Color(), 0, 0); }

    // ---- shapes ----
    Obj push_node(HittablePtr h, int kind, int32_t mat, std::initializer_list<double> f, int32_t child = -1, int32_t axis = 0) {
        RtNode n;
        std::memset(&n, 0, sizeof(n));
        n.kind = kind; n.material = mat; n.first_child = child; n.child_count = 0; n.axis = axis;
        int i = 0;
        for (double v : f) n.f[i++] = v;
        w.nodes.push_back(n);
        w.by_node.push_back(h);
        h->desc_node = (int32_t)w.nodes.size() - 1;
        return Obj{h, h->desc_node};
    }
    Obj sphere(const Point3& c, double r, const Mat& m) {
        return push_node(std::make_shared<Sphere>(c, r, m.m), RT_NODE_SPHERE, m.index, {c.x(), c.y(), c.z(), r});
    }
    Obj xy_rect_(double x0, double x1, double y0, double y1, double z, const Mat& m) {
        return push_node(std::make_shared<RectShape>(xy_rect(x0, x1, y0, y1, z), m.m), RT_NODE_XYRECT, m.index, {x0, x1, y0, y1, z});
    }
    Obj xz_rect_(double x0, double x1, double z0, double z1, double y, const Mat& m) {
        return push_node(std::make_shared<RectShape>(xz_rect(x0, x1, z0, z1, y), m.m), RT_NODE_XZRECT, m.index, {x0, x1, z0, z1, y});
    }
    Obj yz_rect_(double y0, double y1, double z0, double z1, double x, const Mat& m) {
        return push_node(std::make_shared<RectShape>(yz_rect(y0, y1, z0, z1, x), m.m), RT_NODE_YZRECT, m.index, {y0, y1, z0, z1, x});
    }
    Obj block(const Point3& p0, const Point3& p1, const Mat& m) {
        return push_node(std::make_shared<Block>(p0, p1, m.m), RT_NODE_BLOCK, m.index, {p0.x(), p0.y(), p0.z(), p1.x(), p1.y(), p1.z()});
    }
    Obj translate(const Vec3& off, const Obj& o) {
        return push_node(std::make_shared<Translate>(off, o.h), RT_NODE_TRANSLATE, -1, {off.x(), off.y(), off.z()}, o.node);
    }
    Obj rotate(int axis, double angle, const Obj& o) {
        return push_node(std::make_shared<Rotate>(axis, angle, o.h), RT_NODE_ROTATE, -1, {angle}, o.node, axis);
    }
    Obj medium(const Obj& boundary, double d, const Color& c) {  // ConstantMedium::from_color
        Tex t = solid(c);
        Mat iso = isotropic(t);
        return push_node(std::make_shared<ConstantMedium>(boundary.h, d, iso.m), RT_NODE_MEDIUM, iso.index, {d}, boundary.node);
    }
    Obj group(HittablePtr h, int kind, const std::vector<Obj>& items) {
        RtNode n;
        std::memset(&n, 0, sizeof(n));
        n.kind = kind; n.material = -1;
        n.first_child = (int32_t)w.children.size();
        n.child_count = (int32_t)items.size();
        for (const auto& o : items) w.children.push_back(o.node);
        w.nodes.push_back(n);
        w.by_node.push_back(h);
        h->desc_node = (int32_t)w.nodes.size() - 1;
        return Obj{h, h->desc_node};
    }
    Obj bvh(const std::vector<Obj>& items) {  // bhv::BHV::new(&mut builder, rng)
        std::vector<HittablePtr> shapes;
        for (const auto& o : items) shapes.push_back(o.h);
        auto b = std::make_shared<BHV>(shapes, rng);
        w.bvh_axes.push_back(b->axes);
        return group(b, RT_NODE_BVH, items);
    }
    Obj list(const std::vector<Obj>& items) {  // HittableList
        auto l = std::make_shared<HittableList>();
        for (const auto& o : items) l->contents.push_back(o.h);
        return group(l, RT_NODE_LIST, items);
    }
    double rnd01() { return rng.gen_range_f64(0.0, 1.0); }
};

inline void set_camera(WorldOut& w, Vec3 from, Vec3 at, double fov) {
    w.lookfrom = from; w.lookat = at; w.vfov = fov;
}

// worlds.rs:27-60
inline void build_simple(Recipe& R) {
    R.w.background_kind = RT_BG_GRADIENT;
    set_camera(R.w, Vec3(-2, 2, 1), Vec3(0, 0, -1), 20.0);
    Mat ground = R.lambertian(R.solid(0.8, 0.8, 0.0));
    Mat center = R.lambertian(R.solid(0.1, 0.3, 0.5));
    Mat left = R.dielectric(1.5);
    Mat right = R.metal(Color(0.8, 0.6, 0.2), 0.0);
    std::vector<Obj> items;
    items.push_back(R.sphere(Point3(0.0, -100.5, -1.0), 100.0, ground));
    items.push_back(R.sphere(Point3(0.0, 0.0, -1.0), 0.5, center));
    items.push_back(R.sphere(Point3(-1.0, 0.0, -1.0), 0.5, left));
    items.push_back(R.sphere(Point3(-1.0, 0.0, -1.0), -0.4, left));
    items.push_back(R.sphere(Point3(1.0, 0.0, -1.0), 0.5, right));
    R.w.root = R.bvh(items).h;
}

// worlds.rs:79-112 (random) and :128-162 (random_chk): same recipe, the ground texture differs
inline void build_random(Recipe& R, bool checker_ground) {
    R.w.background_kind = RT_BG_GRADIENT;
    set_camera(R.w, Vec3(13, 2, 3), Vec3(0, 0, 0), 20.0);
    std::vector<Obj> items;
    Mat ground;
    if (checker_ground) {
        Tex odd = R.solid(0.2, 0.3, 0.1);
        Tex even = R.solid(0.9, 0.9, 0.9);
        ground = R.lambertian(R.checker(odd, even));
    } else {
        ground = R.lambertian(R.solid(0.5, 0.5, 0.5));
    }
    items.push_back(R.sphere(Point3(0.0, -1000.0, 0.0), 1000.0, ground));
    for (int a = -11; a < 11; a++) {
        for (int b = -11; b < 11; b++) {
            double choose = R.rnd01();
            double cx = (double)a + 0.9 * R.rnd01();
            double cz = (double)b + 0.9 * R.rnd01();
            Point3 center(cx, 0.2, cz);
            if ((center - Point3(4.0, 0.2, 0.0)).length() > 0.9) {
                if (choose < 0.8) {
                    Color c1 = random_vec(0.0, 1.0, R.rng);
                    Color c2 = random_vec(0.0, 1.0, R.rng);
                    Color albedo = c1 * c2;
                    items.push_back(R.sphere(center, 0.2, R.lambertian(R.solid(albedo))));
                } else if (choose < 0.95) {
                    Color albedo = random_vec(0.5, 1.0, R.rng);
                    double fuzz = R.rng.gen_range_f64(0.0, 0.5);
                    items.push_back(R.sphere(center, 0.2, R.metal(albedo, fuzz)));
                } else {
                    items.push_back(R.sphere(center, 0.2, R.dielectric(1.5)));
                }
            }
        }
    }
    items.push_back(R.sphere(Point3(0.0, 1.0, 0.0), 1.0, R.dielectric(1.5)));
    items.push_back(R.sphere(Point3(-4.0, 1.0, 0.0), 1.0, R.lambertian(R.solid(0.4, 0.2, 0.1))));
    items.push_back(R.sphere(Point3(4.0, 1.0, 0.0), 1.0, R.metal(Color(0.7, 0.6, 0.5), 0.0)));
    R.w.root = R.bvh(items).h;
}

// worlds.rs:179-186 — a bare Sphere is the world
inline void build_earth(Recipe& R, const uint8_t* rgb, int iw, int ih) {
    R.w.background_kind = RT_BG_GRADIENT;
    set_camera(R.w, Vec3(13, 2, 3), Vec3(0, 0, 0), 20.0);
    Mat surface = R.lambertian(R.image(rgb, iw, ih));
    R.w.root = R.sphere(Point3(0, 0, 0), 2.0, surface).h;
}

// worlds.rs:203-210 and :227-239
inline void build_two_spheres(Recipe& R, bool lights) {
    R.w.background_kind = lights ? RT_BG_BLACK : RT_BG_GRADIENT;
    if (lights) set_camera(R.w, Vec3(20, 3, 6), Vec3(0, 2, 0), 20.0);
    else set_camera(R.w, Vec3(13, 2, 3), Vec3(0, 0, 0), 20.0);
    Tex pertext = R.noise(4.0);
    std::vector<Obj> items;
    items.push_back(R.sphere(Point3(0.0, -1000.0, 0.0), 1000.0, R.lambertian(pertext)));
    items.push_back(R.sphere(Point3(0.0, 2.0, 0.0), 2.0, R.lambertian(pertext)));
    if (lights) {
        Mat l1 = R.diffuse_light(R.solid(0.0, 7.0, 0.0));
        items.push_back(R.xy_rect_(3.0, 5.0, 1.0, 3.0, -2.0, l1));
        Mat l2 = R.diffuse_light(R.solid(7.0, 0.0, 0.0));
        items.push_back(R.sphere(Point3(0.0, 6.0, 0.0), 1.5, l2));
    }
    R.w.root = R.list(items).h;
}

// worlds.rs:259-287 (cornell_box) and :308-335 (cornell_smoke)
inline void build_cornell(Recipe& R, bool smoke) {
    R.w.background_kind = RT_BG_BLACK;
    set_camera(R.w, Vec3(278, 278, -800), Vec3(278, 278, 0), 40.0);
    Mat red = R.lambertian(R.solid(0.65, 0.05, 0.05));
    Mat white = R.lambertian(R.solid(0.73, 0.73, 0.73));
    Mat green = R.lambertian(R.solid(0.12, 0.45, 0.15));
    Mat light = R.diffuse_light(R.solid(7.0, 7.0, 7.0));
    std::vector<Obj> items;
    items.push_back(R.yz_rect_(0.0, 555.0, 0.0, 555.0, 555.0, green));
    items.push_back(R.yz_rect_(0.0, 555.0, 0.0, 555.0, 0.0, red));
    items.push_back(R.xz_rect_(113.0, 443.0, 127.0, 432.0, 554.0, light));
    items.push_back(R.xz_rect_(0.0, 555.0, 0.0, 555.0, 0.0, white));
    items.push_back(R.xz_rect_(0.0, 555.0, 0.0, 555.0, 555.0, white));
    items.push_back(R.xy_rect_(0.0, 555.0, 0.0, 555.0, 555.0, white));

    Obj large = R.block(Point3(0, 0, 0), Point3(165.0, 330.0, 165.0), white);
    large = R.rotate(1, 15.0, large);
    large = R.translate(Vec3(265.0, 0.0, 295.0), large);
    if (smoke) items.push_back(R.medium(large, 0.01, Color(0, 0, 0)));
    else items.push_back(large);

    Obj small = R.block(Point3(0, 0, 0), Point3(165.0, 165.0, 165.0), white);
    small = R.rotate(1, -18.0, small);
    small = R.translate(Vec3(130.0, 0.0, 65.0), small);
    if (smoke) items.push_back(R.medium(small, 0.01, Color(1, 1, 1)));
    else items.push_back(small);
    R.w.root = R.list(items).h;
}

// worlds.rs:355-365
inline void build_debug_perlin(Recipe& R) {
    R.w.background_kind = RT_BG_GRADIENT;
    set_camera(R.w, Vec3(278, 278, -600), Vec3(278, 278, 0), 40.0);
    std::vector<Obj> items;
    items.push_back(R.sphere(Point3(278.0, 278.0, 0.0), 80.0, R.lambertian(R.noise(0.1))));
    R.w.root = R.list(items).h;
}

// worlds.rs:386-468
inline void build_final_scene(Recipe& R, const uint8_t* rgb, int iw, int ih) {
    R.w.background_kind = RT_BG_BLACK;
    set_camera(R.w, Vec3(478, 278, -600), Vec3(278, 278, 0), 40.0);
    std::vector<Obj> items;
    {  // light
        Mat light = R.diffuse_light(R.solid(9.0, 9.0, 9.0));
        items.push_back(R.xz_rect_(123.0, 423.0, 147.0, 412.0, 554.0, light));
    }
    {  // ground: 20 x 20 blocks of random height in their own BVH
        Mat ground = R.lambertian(R.solid(0.48, 0.83, 0.53));
        std::vector<Obj> blocks;
        for (int i = 0; i < 20; i++)
            for (int j = 0; j < 20; j++) {
                double wd = 100.0;
                double x0 = -1000.0 + (double)i * wd;
                double z0 = -1000.0 + (double)j * wd;
                double y0 = 0.0;
                double x1 = x0 + wd;
                double y1 = R.rng.gen_range_f64(1.0, 70.0);
                double z1 = z0 + wd;
                blocks.push_back(R.block(Point3(x0, y0, z0), Point3(x1, y1, z1), ground));
            }
        items.push_back(R.bvh(blocks));
    }
    items.push_back(R.sphere(Point3(400.0, 400.0, 400.0), 50.0, R.lambertian(R.solid(0.7, 0.3, 0.1))));
    items.push_back(R.sphere(Point3(260.0, 150.0, 45.0), 50.0, R.dielectric(1.5)));
    items.push_back(R.sphere(Point3(0.0, 150.0, 145.0), 50.0, R.metal(Color(0.8, 0.8, 0.9), 1.0)));
    {  // glass sphere filled with blue-ish smoke: the boundary is in the list AND bounds the medium
        Obj boundary = R.sphere(Point3(360.0, 150.0, 145.0), 70.0, R.dielectric(1.5));
        items.push_back(boundary);
        items.push_back(R.medium(boundary, 0.2, Color(0.2, 0.4, 0.9)));
    }
    {  // global fog
        Obj boundary = R.sphere(Point3(0, 0, 0), 1000.0, R.dielectric(1.5));
        items.push_back(R.medium(boundary, 0.0001, Color(1, 1, 1)));
    }
    {  // earth
        Mat surface = R.lambertian(R.image(rgb, iw, ih));
        items.push_back(R.sphere(Point3(400.0, 200.0, 400.0), 100.0, surface));
    }
    items.push_back(R.sphere(Point3(220.0, 280.0, 300.0), 80.0, R.lambertian(R.noise(0.1))));
    {  // foam: 1000 spheres in a BVH, rotated then translated
        std::vector<Obj> foam;
        Mat white = R.lambertian(R.solid(0.73, 0.73, 0.73));
        for (int i = 0; i < 1000; i++) {
            Point3 c = random_vec(0.0, 165.0, R.rng);
            foam.push_back(R.sphere(c, 10.0, white));
        }
        Obj b = R.bvh(foam);
        items.push_back(R.translate(Vec3(-100.0, 270.0, 395.0), R.rotate(1, 15.0, b)));
    }
    R.w.root = R.list(items).h;
}

inline const char* const* world_names(int* n) {
    // registry order of worlds.rs:471-484
    static const char* names[] = {"simple", "random", "random_chk", "two_spheres", "simple_light",
                                  "cornell_box", "cornell_smoke", "earth", "debug_perlin", "final_scene"};
    *n = 10;
    return names;
}

// World::build with rng = Pcg64::seed_from_u64(seed) (main.rs:185).  Returns false for an unknown name or a
// missing earth image (the reference unwrap()-panics there, worlds.rs:180,441).
inline bool build_world(const std::string& name, uint64_t seed, const uint8_t* earth, int ew, int eh, WorldOut& out) {
    Pcg64 rng = Pcg64::seed_from_u64(seed);
    Recipe R(out, rng);
    bool needs_earth = name == "earth" || name == "final_scene";
    if (needs_earth && (!earth || ew <= 0 || eh <= 0)) return false;
    if (name == "simple") build_simple(R);
    else if (name == "random") build_random(R, false);
    else if (name == "random_chk") build_random(R, true);
    else if (name == "two_spheres") build_two_spheres(R, false);
    else if (name == "simple_light") build_two_spheres(R, true);
    else if (name == "cornell_box") build_cornell(R, false);
    else if (name == "cornell_smoke") build_cornell(R, true);
    else if (name == "earth") build_earth(R, earth, ew, eh);
    else if (name == "debug_perlin") build_debug_perlin(R);
    else if (name == "final_scene") build_final_scene(R, earth, ew, eh);
    else return false;
    out.root_node = out.root->desc_node;
    out.draws = rng.draws;
    out.finish();
    return true;
}

// Rebuild an oracle tree from a description (for ad-hoc parity scenes built by the tests).  BVH nodes draw
// their split axes from `rng`; closest-hit results do not depend on them.
inline HittablePtr tree_from_desc(const RtSceneDesc& d, int32_t node, Pcg64& rng,
                                  std::vector<MaterialPtr>& mats, std::vector<std::shared_ptr<Texture>>& texs,
                                  std::vector<HittablePtr>& memo);

inline std::shared_ptr<Texture> texture_from_desc(const RtSceneDesc& d, int32_t ti, std::vector<std::shared_ptr<Texture>>& texs) {
    if (texs[ti]) return texs[ti];
    const RtTexture& t = d.textures[ti];
    std::shared_ptr<Texture> r;
    switch (t.kind) {
        case RT_TEX_SOLID: r = std::make_shared<SolidColor>(Color(t.color[0], t.color[1], t.color[2])); break;
        case RT_TEX_CHECKER: r = std::make_shared<Checker>(texture_from_desc(d, t.a, texs), texture_from_desc(d, t.b, texs)); break;
        case RT_TEX_NOISE: {
            auto p = std::make_shared<Perlin>();
            const RtPerlin& rp = d.perlins[t.a];
            for (int i = 0; i < Perlin::N; i++) {
                p->ranvec[i] = Vec3(rp.ranvec[i][0], rp.ranvec[i][1], rp.ranvec[i][2]);
                p->perm_x[i] = rp.perm_x[i]; p->perm_y[i] = rp.perm_y[i]; p->perm_z[i] = rp.perm_z[i];
            }
            r = std::make_shared<NoiseTexture>(p, t.scale);
            break;
        }
        case RT_TEX_IMAGE: {
            const RtImage& im = d.images[t.a];
            auto data = std::make_shared<std::vector<uint8_t>>(im.rgb, im.rgb + (size_t)3 * im.width * im.height);
            r = std::make_shared<ImageTexture>(im.width, im.height, data);
            break;
        }
    }
    texs[ti] = r;
    return r;
}

inline MaterialPtr material_from_desc(const RtSceneDesc& d, int32_t mi, std::vector<MaterialPtr>& mats,
                                      std::vector<std::shared_ptr<Texture>>& texs) {
    if (mi < 0) return nullptr;
    if (mats[mi]) return mats[mi];
    const RtMaterial& m = d.materials[mi];
    MaterialPtr r;
    switch (m.kind) {
        case RT_MAT_LAMBERTIAN: r = std::make_shared<Lambertian>(texture_from_desc(d, m.texture, texs)); break;
        case RT_MAT_METAL: r = std::make_shared<Metal>(Color(m.albedo[0], m.albedo[1], m.albedo[2]), m.fuzz); break;
        case RT_MAT_DIELECTRIC: r = std::make_shared<Dielectric>(m.ior); break;
        case RT_MAT_DIFFUSE_LIGHT: r = std::make_shared<DiffuseLight>(texture_from_desc(d, m.texture, texs)); break;
        case RT_MAT_ISOTROPIC: r = std::make_shared<Isotropic>(texture_from_desc(d, m.texture, texs)); break;
    }
    r->desc_index = mi;
    mats[mi] = r;
    return r;
}

inline HittablePtr tree_from_desc(const RtSceneDesc& d, int32_t node, Pcg64& rng, std::vector<MaterialPtr>& mats,
                                  std::vector<std::shared_ptr<Texture>>& texs, std::vector<HittablePtr>& memo) {
    if (memo[node]) return memo[node];
    const RtNode& n = d.nodes[node];
    const double* f = n.f;
    MaterialPtr m = material_from_desc(d, n.material, mats, texs);
    HittablePtr r;
    switch (n.kind) {
        case RT_NODE_SPHERE: r = std::make_shared<Sphere>(Point3(f[0], f[1], f[2]), f[3], m); break;
        case RT_NODE_MOVING_SPHERE:  // extension, see scene.hpp
            r = std::make_shared<MovingSphere>(Point3(f[0], f[1], f[2]), Point3(f[3], f[4], f[5]), f[6], m);
            break;
        case RT_NODE_XYRECT: r = std::make_shared<RectShape>(xy_rect(f[0], f[1], f[2], f[3], f[4]), m); break;
        case RT_NODE_XZRECT: r = std::make_shared<RectShape>(xz_rect(f[0], f[1], f[2], f[3], f[4]), m); break;
        case RT_NODE_YZRECT: r = std::make_shared<RectShape>(yz_rect(f[0], f[1], f[2], f[3], f[4]), m); break;
        case RT_NODE_BLOCK: r = std::make_shared<Block>(Point3(f[0], f[1], f[2]), Point3(f[3], f[4], f[5]), m); break;
        case RT_NODE_TRANSLATE:
            r = std::make_shared<Translate>(Vec3(f[0], f[1], f[2]), tree_from_desc(d, n.first_child, rng, mats, texs, memo));
            break;
        case RT_NODE_ROTATE:
            r = std::make_shared<Rotate>(n.axis, f[0], tree_from_desc(d, n.first_child, rng, mats, texs, memo));
            break;
        case RT_NODE_MEDIUM:
            r = std::make_shared<ConstantMedium>(tree_from_desc(d, n.first_child, rng, mats, texs, memo), f[0], m);
            break;
        case RT_NODE_BVH: {
            std::vector<HittablePtr> shapes;
            for (int i = 0; i < n.child_count; i++)
                shapes.push_back(tree_from_desc(d, d.children[n.first_child + i], rng, mats, texs, memo));
            r = std::make_shared<BHV>(shapes, rng);
            break;
        }
        case RT_NODE_LIST: {
            auto l = std::make_shared<HittableList>();
            for (int i = 0; i < n.child_count; i++)
                l->contents.push_back(tree_from_desc(d, d.children[n.first_child + i], rng, mats, texs, memo));
            r = l;
            break;
        }
        default: return nullptr;
    }
    r->desc_node = node;
    memo[node] = r;
    return r;
}

}  // namespace orc
