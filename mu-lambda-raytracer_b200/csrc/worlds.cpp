// Host side above the C ABI: the reference's scene recipes (src/worlds.rs) restated so that they EMIT a
// scene description instead of an opaque Box<dyn Hittable> (src/worlds.rs:18 returns a trait object that
// cannot be enumerated, so "build then flatten" has to start from a description).
//
// Emission rule (shared with the test oracle so the two descriptions hash identically): an entry is
// appended where the reference calls the constructor; `.clone()` re-uses the entry.  The world RNG is
// consumed in exactly the reference's order, including the one `gen_range(0..3)` per inner node that
// `bhv::Node::new` draws while building a BVH (src/bhv.rs:127) — the device builds its own BVH, but the
// draws must still happen to keep everything built afterwards on the same stream.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rt_b200.h"
#include "host_rng.h"
#include "internal.h"

namespace rtb {


namespace {

struct V3 {
    double x, y, z;
};

class Sink {
   public:
    Sink(OwnedDesc& out, WorldRng& rng) : o(out), rng(rng) {}
    OwnedDesc& o;
    WorldRng& rng;

    int solid(double r, double g, double b) {  // textures.rs:12-20
        RtTexture t{};
        t.kind = RT_TEX_SOLID;
        t.a = t.b = -1;
        t.color[0] = r, t.color[1] = g, t.color[2] = b;
        o.textures.push_back(t);
        return (int)o.textures.size() - 1;
    }
    int checker(int odd, int even) {  // textures.rs:34-38
        RtTexture t{};
        t.kind = RT_TEX_CHECKER;
        t.a = odd, t.b = even;
        o.textures.push_back(t);
        return (int)o.textures.size() - 1;
    }
    // NoiseTexture::new -> Perlin::new (textures.rs:62-74, :136-148): 1024 unit gradients, then three
    // permutations shuffled from the top with gen_range(0..i) (i exclusive)
    int noise(double scale) {
        o.perlins.emplace_back();
        RtPerlin& p = o.perlins.back();
        for (int i = 0; i < RT_PERLIN_POINTS; ++i) {
            double a = rng.range_f64(-1.0, 1.0), b = rng.range_f64(-1.0, 1.0), c = rng.range_f64(-1.0, 1.0);
            double len = std::sqrt(a * a + b * b + c * c);
            p.ranvec[i][0] = a / len, p.ranvec[i][1] = b / len, p.ranvec[i][2] = c / len;
        }
        int32_t* perms[3] = {p.perm_x, p.perm_y, p.perm_z};
        for (int32_t* perm : perms) {
            for (int i = 0; i < RT_PERLIN_POINTS; ++i) perm[i] = i;
            for (int i = RT_PERLIN_POINTS - 1; i > 0; --i) {
                uint64_t j = rng.range_usize(0, (uint64_t)i);
                int32_t tmp = perm[i];
                perm[i] = perm[j];
                perm[j] = tmp;
            }
        }
        RtTexture t{};
        t.kind = RT_TEX_NOISE;
        t.a = (int)o.perlins.size() - 1, t.b = -1;
        t.scale = scale;
        o.textures.push_back(t);
        return (int)o.textures.size() - 1;
    }
    int image(const uint8_t* rgb, int w, int h) {  // image_texture.rs:10-14
        o.pixels.emplace_back(rgb, rgb + (size_t)3 * w * h);
        RtImage im{};
        im.width = w, im.height = h;
        im.rgb = nullptr;  // patched in seal(): vectors may still move
        o.images.push_back(im);
        RtTexture t{};
        t.kind = RT_TEX_IMAGE;
        t.a = (int)o.images.size() - 1, t.b = -1;
        o.textures.push_back(t);
        return (int)o.textures.size() - 1;
    }

    int material(int kind, int tex, V3 albedo, double fuzz, double ior) {
        RtMaterial m{};
        m.kind = kind, m.texture = tex;
        m.albedo[0] = albedo.x, m.albedo[1] = albedo.y, m.albedo[2] = albedo.z;
        m.fuzz = fuzz, m.ior = ior;
        o.materials.push_back(m);
        return (int)o.materials.size() - 1;
    }
    int lambertian(int tex) { return material(RT_MAT_LAMBERTIAN, tex, {0, 0, 0}, 0, 0); }
    int metal(V3 albedo, double fuzz) { return material(RT_MAT_METAL, -1, albedo, fuzz, 0); }
    int dielectric(double ior) { return material(RT_MAT_DIELECTRIC, -1, {0, 0, 0}, 0, ior); }
    int light(int tex) { return material(RT_MAT_DIFFUSE_LIGHT, tex, {0, 0, 0}, 0, 0); }

    int node(int kind, int mat, const double* f, int nf, int child = -1, int axis = 0) {
        RtNode n{};
        n.kind = kind, n.material = mat, n.first_child = child, n.axis = axis;
        for (int i = 0; i < nf; ++i) n.f[i] = f[i];
        o.nodes.push_back(n);
        return (int)o.nodes.size() - 1;
    }
    int sphere(V3 c, double r, int mat) {
        double f[4] = {c.x, c.y, c.z, r};
        return node(RT_NODE_SPHERE, mat, f, 4);
    }
    int rect(int kind, double a0, double a1, double b0, double b1, double k, int mat) {
        double f[5] = {a0, a1, b0, b1, k};
        return node(kind, mat, f, 5);
    }
    int block(V3 p0, V3 p1, int mat) {
        double f[6] = {p0.x, p0.y, p0.z, p1.x, p1.y, p1.z};
        return node(RT_NODE_BLOCK, mat, f, 6);
    }
    int translate(V3 off, int child) {
        double f[3] = {off.x, off.y, off.z};
        return node(RT_NODE_TRANSLATE, -1, f, 3, child);
    }
    int rotate(int axis, double degrees, int child) { return node(RT_NODE_ROTATE, -1, &degrees, 1, child, axis); }
    int medium(int boundary, double density, V3 color) {  // ConstantMedium::from_color (volumes.rs:19-23)
        int tex = solid(color.x, color.y, color.z);
        int iso = material(RT_MAT_ISOTROPIC, tex, {0, 0, 0}, 0, 0);
        return node(RT_NODE_MEDIUM, iso, &density, 1, boundary);
    }
    int group(int kind, const std::vector<int>& items) {
        RtNode n{};
        n.kind = kind, n.material = -1;
        n.first_child = (int)o.children.size();
        n.child_count = (int)items.size();
        o.children.insert(o.children.end(), items.begin(), items.end());
        o.nodes.push_back(n);
        return (int)o.nodes.size() - 1;
    }
    int list(const std::vector<int>& items) { return group(RT_NODE_LIST, items); }
    // bhv::BHV::new: Node::new splits len/2 and draws one axis per inner node, pre-order (bhv.rs:122-145).
    // The split sizes depend only on the count, so the draws can be replayed without sorting anything.
    void draw_bvh_axes(size_t count) {
        if (count < 2) return;
        (void)rng.range_usize(0, 3);
        draw_bvh_axes(count / 2);
        draw_bvh_axes(count - count / 2);
    }
    int bvh(const std::vector<int>& items) {
        draw_bvh_axes(items.size());
        return group(RT_NODE_BVH, items);
    }
    V3 random_v3(double lo, double hi) {  // Vec3::random (vec.rs:15-17): x, y, z in this order
        double a = rng.range_f64(lo, hi), b = rng.range_f64(lo, hi), c = rng.range_f64(lo, hi);
        return {a, b, c};
    }
};

void camera(RtWorldInfo& wi, V3 from, V3 at, double fov, int bg) {
    wi.lookfrom[0] = from.x, wi.lookfrom[1] = from.y, wi.lookfrom[2] = from.z;
    wi.lookat[0] = at.x, wi.lookat[1] = at.y, wi.lookat[2] = at.z;
    wi.vfov_deg = fov;
    wi.background_kind = bg;
}

// ---- the ten recipes; each returns the root node ----

int world_simple(Sink& s) {  // worlds.rs:41-59
    int ground = s.lambertian(s.solid(0.8, 0.8, 0.0));
    int center = s.lambertian(s.solid(0.1, 0.3, 0.5));
    int left = s.dielectric(1.5);
    int right = s.metal({0.8, 0.6, 0.2}, 0.0);
    std::vector<int> items = {
        s.sphere({0.0, -100.5, -1.0}, 100.0, ground), s.sphere({0.0, 0.0, -1.0}, 0.5, center),
        s.sphere({-1.0, 0.0, -1.0}, 0.5, left),      s.sphere({-1.0, 0.0, -1.0}, -0.4, left),
        s.sphere({1.0, 0.0, -1.0}, 0.5, right),
    };
    return s.bvh(items);
}

int world_random(Sink& s, bool checker_ground) {  // worlds.rs:79-112, :128-162
    std::vector<int> items;
    int ground_tex;
    if (checker_ground) {
        int odd = s.solid(0.2, 0.3, 0.1);
        int even = s.solid(0.9, 0.9, 0.9);
        ground_tex = s.checker(odd, even);
    } else {
        ground_tex = s.solid(0.5, 0.5, 0.5);
    }
    items.push_back(s.sphere({0.0, -1000.0, 0.0}, 1000.0, s.lambertian(ground_tex)));
    for (int a = -11; a < 11; ++a) {
        for (int b = -11; b < 11; ++b) {
            double choose = s.rng.unit();
            double cx = (double)a + 0.9 * s.rng.unit();
            double cz = (double)b + 0.9 * s.rng.unit();
            double dx = cx - 4.0, dy = 0.2 - 0.2, dz = cz - 0.0;
            if (std::sqrt(dx * dx + dy * dy + dz * dz) > 0.9) {
                if (choose < 0.8) {
                    V3 p = s.random_v3(0.0, 1.0), q = s.random_v3(0.0, 1.0);
                    int tex = s.solid(p.x * q.x, p.y * q.y, p.z * q.z);
                    items.push_back(s.sphere({cx, 0.2, cz}, 0.2, s.lambertian(tex)));
                } else if (choose < 0.95) {
                    V3 albedo = s.random_v3(0.5, 1.0);
                    double fuzz = s.rng.range_f64(0.0, 0.5);
                    items.push_back(s.sphere({cx, 0.2, cz}, 0.2, s.metal(albedo, fuzz)));
                } else {
                    items.push_back(s.sphere({cx, 0.2, cz}, 0.2, s.dielectric(1.5)));
                }
            }
        }
    }
    items.push_back(s.sphere({0.0, 1.0, 0.0}, 1.0, s.dielectric(1.5)));
    items.push_back(s.sphere({-4.0, 1.0, 0.0}, 1.0, s.lambertian(s.solid(0.4, 0.2, 0.1))));
    items.push_back(s.sphere({4.0, 1.0, 0.0}, 1.0, s.metal({0.7, 0.6, 0.5}, 0.0)));
    return s.bvh(items);
}

int world_earth(Sink& s, const uint8_t* rgb, int w, int h) {  // worlds.rs:179-186
    return s.sphere({0, 0, 0}, 2.0, s.lambertian(s.image(rgb, w, h)));
}

int world_two_spheres(Sink& s, bool lights) {  // worlds.rs:203-210, :227-239
    int pertext = s.noise(4.0);
    std::vector<int> items;
    items.push_back(s.sphere({0.0, -1000.0, 0.0}, 1000.0, s.lambertian(pertext)));
    items.push_back(s.sphere({0.0, 2.0, 0.0}, 2.0, s.lambertian(pertext)));
    if (lights) {
        items.push_back(s.rect(RT_NODE_XYRECT, 3.0, 5.0, 1.0, 3.0, -2.0, s.light(s.solid(0.0, 7.0, 0.0))));
        items.push_back(s.sphere({0.0, 6.0, 0.0}, 1.5, s.light(s.solid(7.0, 0.0, 0.0))));
    }
    return s.list(items);
}

int world_cornell(Sink& s, bool smoke) {  // worlds.rs:259-287, :308-335
    int red = s.lambertian(s.solid(0.65, 0.05, 0.05));
    int white = s.lambertian(s.solid(0.73, 0.73, 0.73));
    int green = s.lambertian(s.solid(0.12, 0.45, 0.15));
    int lamp = s.light(s.solid(7.0, 7.0, 7.0));
    std::vector<int> items;
    items.push_back(s.rect(RT_NODE_YZRECT, 0.0, 555.0, 0.0, 555.0, 555.0, green));
    items.push_back(s.rect(RT_NODE_YZRECT, 0.0, 555.0, 0.0, 555.0, 0.0, red));
    items.push_back(s.rect(RT_NODE_XZRECT, 113.0, 443.0, 127.0, 432.0, 554.0, lamp));
    items.push_back(s.rect(RT_NODE_XZRECT, 0.0, 555.0, 0.0, 555.0, 0.0, white));
    items.push_back(s.rect(RT_NODE_XZRECT, 0.0, 555.0, 0.0, 555.0, 555.0, white));
    items.push_back(s.rect(RT_NODE_XYRECT, 0.0, 555.0, 0.0, 555.0, 555.0, white));
    int large = s.translate({265.0, 0.0, 295.0}, s.rotate(1, 15.0, s.block({0, 0, 0}, {165.0, 330.0, 165.0}, white)));
    items.push_back(smoke ? s.medium(large, 0.01, {0, 0, 0}) : large);
    int small = s.translate({130.0, 0.0, 65.0}, s.rotate(1, -18.0, s.block({0, 0, 0}, {165.0, 165.0, 165.0}, white)));
    items.push_back(smoke ? s.medium(small, 0.01, {1, 1, 1}) : small);
    return s.list(items);
}

int world_debug_perlin(Sink& s) {  // worlds.rs:355-365
    std::vector<int> items = {s.sphere({278.0, 278.0, 0.0}, 80.0, s.lambertian(s.noise(0.1)))};
    return s.list(items);
}

int world_final_scene(Sink& s, const uint8_t* rgb, int w, int h) {  // worlds.rs:386-468
    std::vector<int> items;
    items.push_back(s.rect(RT_NODE_XZRECT, 123.0, 423.0, 147.0, 412.0, 554.0, s.light(s.solid(9.0, 9.0, 9.0))));
    {
        int ground = s.lambertian(s.solid(0.48, 0.83, 0.53));
        std::vector<int> blocks;
        for (int i = 0; i < 20; ++i)
            for (int j = 0; j < 20; ++j) {
                const double side = 100.0;
                double x0 = -1000.0 + (double)i * side, z0 = -1000.0 + (double)j * side;
                double y1 = s.rng.range_f64(1.0, 70.0);
                blocks.push_back(s.block({x0, 0.0, z0}, {x0 + side, y1, z0 + side}, ground));
            }
        items.push_back(s.bvh(blocks));
    }
    items.push_back(s.sphere({400.0, 400.0, 400.0}, 50.0, s.lambertian(s.solid(0.7, 0.3, 0.1))));
    items.push_back(s.sphere({260.0, 150.0, 45.0}, 50.0, s.dielectric(1.5)));
    items.push_back(s.sphere({0.0, 150.0, 145.0}, 50.0, s.metal({0.8, 0.8, 0.9}, 1.0)));
    {
        int shell = s.sphere({360.0, 150.0, 145.0}, 70.0, s.dielectric(1.5));
        items.push_back(shell);
        items.push_back(s.medium(shell, 0.2, {0.2, 0.4, 0.9}));
    }
    items.push_back(s.medium(s.sphere({0, 0, 0}, 1000.0, s.dielectric(1.5)), 0.0001, {1, 1, 1}));
    items.push_back(s.sphere({400.0, 200.0, 400.0}, 100.0, s.lambertian(s.image(rgb, w, h))));
    items.push_back(s.sphere({220.0, 280.0, 300.0}, 80.0, s.lambertian(s.noise(0.1))));
    {
        int white = s.lambertian(s.solid(0.73, 0.73, 0.73));
        std::vector<int> foam;
        for (int k = 0; k < 1000; ++k) foam.push_back(s.sphere(s.random_v3(0.0, 165.0), 10.0, white));
        items.push_back(s.translate({-100.0, 270.0, 395.0}, s.rotate(1, 15.0, s.bvh(foam))));
    }
    return s.list(items);
}

struct Entry {
    const char* name;
    V3 from, at;
    double fov;
    int bg;
    bool earth, rng;
};
// registry order and default cameras/backgrounds of worlds.rs:27-40, 66-77, ..., 471-484
const Entry kWorlds[] = {
    {"simple", {-2, 2, 1}, {0, 0, -1}, 20.0, RT_BG_GRADIENT, false, true},
    {"random", {13, 2, 3}, {0, 0, 0}, 20.0, RT_BG_GRADIENT, false, true},
    {"random_chk", {13, 2, 3}, {0, 0, 0}, 20.0, RT_BG_GRADIENT, false, true},
    {"two_spheres", {13, 2, 3}, {0, 0, 0}, 20.0, RT_BG_GRADIENT, false, true},
    {"simple_light", {20, 3, 6}, {0, 2, 0}, 20.0, RT_BG_BLACK, false, true},
    {"cornell_box", {278, 278, -800}, {278, 278, 0}, 40.0, RT_BG_BLACK, false, false},
    {"cornell_smoke", {278, 278, -800}, {278, 278, 0}, 40.0, RT_BG_BLACK, false, false},
    {"earth", {13, 2, 3}, {0, 0, 0}, 20.0, RT_BG_GRADIENT, true, false},
    {"debug_perlin", {278, 278, -600}, {278, 278, 0}, 40.0, RT_BG_GRADIENT, false, true},
    {"final_scene", {478, 278, -600}, {278, 278, 0}, 40.0, RT_BG_BLACK, true, true},
};
const int kWorldCount = sizeof(kWorlds) / sizeof(kWorlds[0]);

const Entry* find_world(const char* name) {
    if (!name) return nullptr;
    for (const Entry& e : kWorlds)
        if (std::strcmp(e.name, name) == 0) return &e;
    return nullptr;
}

}  // namespace

void seal_desc(OwnedDesc& o, int root, int background_kind) {
    RtSceneDesc& d = o.d;
    std::memset(&d, 0, sizeof d);
    d.root = root;
    d.background_kind = background_kind;
    if (background_kind == RT_BG_GRADIENT) {  // GradientBackground::default() (raytrace.rs:21-26)
        d.background_top[0] = 0.5, d.background_top[1] = 0.7, d.background_top[2] = 1.0;
        d.background_bottom[0] = d.background_bottom[1] = d.background_bottom[2] = 1.0;
    }
    for (size_t i = 0; i < o.images.size(); ++i) o.images[i].rgb = o.pixels[i].data();
    d.n_nodes = (int32_t)o.nodes.size(), d.nodes = o.nodes.data();
    d.n_children = (int32_t)o.children.size(), d.children = o.children.data();
    d.n_materials = (int32_t)o.materials.size(), d.materials = o.materials.data();
    d.n_textures = (int32_t)o.textures.size(), d.textures = o.textures.data();
    d.n_perlins = (int32_t)o.perlins.size(), d.perlins = o.perlins.data();
    d.n_images = (int32_t)o.images.size(), d.images = o.images.data();
}

OwnedDesc* clone_desc(const RtSceneDesc* src) {
    OwnedDesc* o = new OwnedDesc();
    o->nodes.assign(src->nodes, src->nodes + src->n_nodes);
    if (src->n_children > 0) o->children.assign(src->children, src->children + src->n_children);
    if (src->n_materials > 0) o->materials.assign(src->materials, src->materials + src->n_materials);
    if (src->n_textures > 0) o->textures.assign(src->textures, src->textures + src->n_textures);
    if (src->n_perlins > 0) o->perlins.assign(src->perlins, src->perlins + src->n_perlins);
    for (int i = 0; i < src->n_images; ++i) {
        const RtImage& im = src->images[i];
        o->images.push_back(im);
        o->pixels.emplace_back(im.rgb, im.rgb + (size_t)3 * im.width * im.height);
    }
    double top[3], bottom[3];
    std::memcpy(top, src->background_top, sizeof top), std::memcpy(bottom, src->background_bottom, sizeof bottom);
    seal_desc(*o, src->root, src->background_kind);
    std::memcpy(o->d.background_top, top, sizeof top), std::memcpy(o->d.background_bottom, bottom, sizeof bottom);
    return o;
}

void free_desc(OwnedDesc* o) { delete o; }

}  // namespace rtb

using namespace rtb;

extern "C" {

int rt_world_count(void) { return kWorldCount; }

const char* rt_world_name(int index) { return index >= 0 && index < kWorldCount ? kWorlds[index].name : nullptr; }

int rt_world_info(const char* name, RtWorldInfo* out) {
    const Entry* e = find_world(name);
    if (!e || !out) return set_error(RT_ERR_INVALID, "unknown world '%s'", name ? name : "(null)");
    camera(*out, e->from, e->at, e->fov, e->bg);
    out->needs_earthmap = e->earth;
    out->uses_rng = e->rng;
    return RT_OK;
}

int rt_world_build(const char* name, uint64_t seed, const uint8_t* earth_rgb, int32_t earth_w, int32_t earth_h,
                   RtSceneDesc** out, uint64_t* n_draws) {
    const Entry* e = find_world(name);
    if (!e || !out) return set_error(RT_ERR_INVALID, "unknown world '%s'", name ? name : "(null)");
    if (e->earth && (!earth_rgb || earth_w <= 0 || earth_h <= 0))
        return set_error(RT_ERR_INVALID, "world '%s' needs the earthmap image (RGB8)", name);
    OwnedDesc* od = new OwnedDesc();
    WorldRng rng(seed);  // rngator.rng(0) = Pcg64::seed_from_u64(seed + 0)  (main.rs:185)
    Sink s(*od, rng);
    std::string n(name);
    int root;
    if (n == "simple") root = world_simple(s);
    else if (n == "random") root = world_random(s, false);
    else if (n == "random_chk") root = world_random(s, true);
    else if (n == "two_spheres") root = world_two_spheres(s, false);
    else if (n == "simple_light") root = world_two_spheres(s, true);
    else if (n == "cornell_box") root = world_cornell(s, false);
    else if (n == "cornell_smoke") root = world_cornell(s, true);
    else if (n == "earth") root = world_earth(s, earth_rgb, earth_w, earth_h);
    else if (n == "debug_perlin") root = world_debug_perlin(s);
    else root = world_final_scene(s, earth_rgb, earth_w, earth_h);
    seal_desc(*od, root, e->bg);
    if (n_draws) *n_draws = rng.calls();
    *out = &od->d;
    return RT_OK;
}

void rt_scene_desc_free(RtSceneDesc* desc) {
    if (desc) delete reinterpret_cast<OwnedDesc*>(desc);
}

}  // extern "C"
