// ORACLE — test infrastructure only.  Nothing under oracle/ is part of the product path.
//
// C entry points over the oracle (pcg64.hpp, scene.hpp, worlds.hpp, render.hpp) so that tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs can drive it through ctypes.
// Built by oracle/Makefile into oracle/liboracle.so with -O3 -ffp-contract=off (Rust never fuses a*b+c).
#include <chrono>
#include <cstdio>

#include "render.hpp"
#include "worlds.hpp"

using namespace orc;

namespace {
struct OrcWorld {
    WorldOut w;
    // for worlds rebuilt from a description
    std::vector<MaterialPtr> mats;
    std::vector<std::shared_ptr<Texture>> texs;
    std::vector<HittablePtr> memo;
    bool from_desc = false;
    RtSceneDesc foreign;  // the caller's description (from_desc)
    HittablePtr node(int32_t n) {
        if (n < 0) return w.root;
        if (from_desc) return (size_t)n < memo.size() ? memo[n] : nullptr;
        return (size_t)n < w.by_node.size() ? w.by_node[n] : nullptr;
    }
};
}  // namespace

extern "C" {

// one closest-hit record of Hittable::hit, all f64 (hittable.rs:7-15)
struct OrcHit {
    int32_t hit;
    int32_t front_face;
    int32_t material;  // description material index
    int32_t node;      // description node of the primitive
    double t, u, v;
    double p[3];
    double normal[3];
};

void* orc_world_build(const char* name, uint64_t seed, const uint8_t* earth, int ew, int eh) {
    OrcWorld* W = new OrcWorld();
    if (!build_world(name, seed, earth, ew, eh, W->w)) {
        delete W;
        return nullptr;
    }
    return W;
}

void* orc_world_from_desc(const RtSceneDesc* d, uint64_t bvh_seed) {
    OrcWorld* W = new OrcWorld();
    W->from_desc = true;
    W->foreign = *d;
    W->mats.resize(d->n_materials);
    W->texs.resize(d->n_textures);
    W->memo.resize(d->n_nodes);
    Pcg64 rng = Pcg64::seed_from_u64(bvh_seed);
    // build every node so that any subtree can be addressed by the parity harness
    for (int32_t n = 0; n < d->n_nodes; n++)
        if (!tree_from_desc(*d, n, rng, W->mats, W->texs, W->memo)) {
            delete W;
            return nullptr;
        }
    W->w.root = W->memo[d->root];
    W->w.root_node = d->root;
    W->w.background_kind = d->background_kind;
    return W;
}

void orc_world_free(void* h) { delete (OrcWorld*)h; }

const RtSceneDesc* orc_world_desc(void* h) {
    OrcWorld* W = (OrcWorld*)h;
    return W->from_desc ? &W->foreign : &W->w.desc;
}

void orc_world_info(void* h, double lookfrom[3], double lookat[3], double* vfov, int32_t* background, uint64_t* draws,
                    int32_t* n_bvh) {
    OrcWorld* W = (OrcWorld*)h;
    for (int i = 0; i < 3; i++) {
        lookfrom[i] = W->w.lookfrom.e[i];
        lookat[i] = W->w.lookat.e[i];
    }
    *vfov = W->w.vfov;
    *background = W->w.background_kind;
    *draws = W->w.draws;
    *n_bvh = (int32_t)W->w.bvh_axes.size();
}

int32_t orc_world_bvh_axes(void* h, int32_t which, int32_t* out, int32_t cap) {
    OrcWorld* W = (OrcWorld*)h;
    if (which < 0 || (size_t)which >= W->w.bvh_axes.size()) return -1;
    const auto& a = W->w.bvh_axes[which];
    for (int32_t i = 0; i < cap && (size_t)i < a.size(); i++) out[i] = a[i];
    return (int32_t)a.size();
}

// Renderer::render (raytrace.rs:172-186).  counters[16]: paths, rays, aabb, sphere, rect, xform, medium,
// scatter[1..5], perlin, image, background, depth_exhausted.  Returns wall seconds of the render region
// (the reference times exactly this region, main.rs:145-174).
double orc_render(void* h, const RtCamera* cam, int32_t width, int32_t height, int32_t spp, int32_t max_depth,
                  uint64_t render_seed, int32_t row_begin, int32_t row_end, int32_t threads, double* accum,
                  int32_t* rgb, uint64_t* counters) {
    OrcWorld* W = (OrcWorld*)h;
    RenderJob J;
    J.world = W->w.root.get();
    const RtSceneDesc* d = orc_world_desc(h);
    J.bg.kind = d->background_kind;
    J.bg.top = Color(d->background_top[0], d->background_top[1], d->background_top[2]);
    J.bg.bottom = Color(d->background_bottom[0], d->background_bottom[1], d->background_bottom[2]);
    J.cam = Camera(Point3(cam->lookfrom[0], cam->lookfrom[1], cam->lookfrom[2]),
                   Point3(cam->lookat[0], cam->lookat[1], cam->lookat[2]), Vec3(cam->vup[0], cam->vup[1], cam->vup[2]),
                   cam->vfov_deg, cam->aspect_ratio, cam->aperture, cam->focus_dist);
    J.cam.time0 = cam->time0, J.cam.time1 = cam->time1;
    J.width = width; J.height = height; J.spp = spp; J.max_depth = max_depth;
    J.render_seed = render_seed;
    J.row_begin = row_begin; J.row_end = row_end;
    Counters c;
    auto t0 = std::chrono::steady_clock::now();
    render(J, threads, accum, rgb, &c);
    auto t1 = std::chrono::steady_clock::now();
    if (counters) {
        uint64_t v[16] = {c.paths, c.rays, c.aabb_tests, c.sphere_tests, c.rect_tests, c.xform, c.medium_tests,
                          c.scatter[1], c.scatter[2], c.scatter[3], c.scatter[4], c.scatter[5],
                          c.perlin_evals, c.image_evals, c.background_evals, c.depth_exhausted};
        for (int i = 0; i < 16; i++) counters[i] = v[i];
    }
    return std::chrono::duration<double>(t1 - t0).count();
}

int32_t orc_hardware_threads() { return (int32_t)std::thread::hardware_concurrency(); }

// Hittable::hit of the subtree at description node `node` (-1 = root) for N rays given as 8 f64 each:
// origin, direction (not normalised), t_min, t_max.  Media draw their free-flight sample from
// Pcg64::seed_from_u64(rng_seed + ray index).
int32_t orc_hit_batch_at(void* h, int32_t node, double time, const double* rays, int64_t n, uint64_t rng_seed, int32_t skip_media, OrcHit* out);
int32_t orc_hit_batch(void* h, int32_t node, const double* rays, int64_t n, uint64_t rng_seed, int32_t skip_media, OrcHit* out) {
    return orc_hit_batch_at(h, node, 0.0, rays, n, rng_seed, skip_media, out);
}
// the same with the rays at time `time` (moving-sphere extension)
int32_t orc_hit_batch_at(void* h, int32_t node, double time, const double* rays, int64_t n, uint64_t rng_seed, int32_t skip_media, OrcHit* out) {
    OrcWorld* W = (OrcWorld*)h;
    HittablePtr obj = W->node(node);
    if (!obj) return -1;
    for (int64_t i = 0; i < n; i++) {
        const double* r = rays + 8 * i;
        Ray ray{Point3(r[0], r[1], r[2]), Vec3(r[3], r[4], r[5]), time};
        Ctx cx;
        cx.rng = Pcg64::seed_from_u64(rng_seed + (uint64_t)i);
        cx.skip_media = skip_media != 0;
        Hit hit;
        OrcHit& o = out[i];
        std::memset(&o, 0, sizeof(o));
        o.material = -1; o.node = -1;
        if (obj->hit(ray, r[6], r[7], cx, hit)) {
            o.hit = 1;
            o.front_face = hit.front_face ? 1 : 0;
            o.material = hit.material ? hit.material->desc_index : -1;
            o.node = hit.desc_node;
            o.t = hit.t; o.u = hit.u; o.v = hit.v;
            for (int k = 0; k < 3; k++) {
                o.p[k] = hit.p.e[k];
                o.normal[k] = hit.normal.e[k];
            }
        }
    }
    return 0;
}

// the deterministic part of ConstantMedium::hit (volumes.rs:27-43): clipped boundary interval
int32_t orc_medium_interval_batch(void* h, int32_t node, const double* rays, int64_t n, int32_t* hit, double* t1t2) {
    OrcWorld* W = (OrcWorld*)h;
    HittablePtr obj = W->node(node);
    ConstantMedium* m = dynamic_cast<ConstantMedium*>(obj.get());
    if (!m) return -1;
    for (int64_t i = 0; i < n; i++) {
        const double* r = rays + 8 * i;
        Ray ray{Point3(r[0], r[1], r[2]), Vec3(r[3], r[4], r[5])};
        Ctx cx;
        double t1 = 0, t2 = 0;
        hit[i] = m->interval(ray, r[6], r[7], cx, t1, t2) ? 1 : 0;
        t1t2[2 * i] = t1;
        t1t2[2 * i + 1] = t2;
    }
    return 0;
}

// Material::scatter / emit (materials.rs:7-11, :25-127; volumes.rs:77-83) of description material `material` for a batch
// of caller-supplied hits.  in: n x 15 doubles (ray origin, ray dir, p, normal, u, v, front_face); item i draws from
// Pcg64::seed_from_u64(seed_base + i).  out: n x 11 doubles (scattered, attenuation rgb, direction xyz, emitted rgb, the
// first unit() of that stream — what a Dielectric compares its reflectance with, materials.rs:98).
int32_t orc_scatter_batch(void* h, int32_t material, const double* in, int64_t n, uint64_t seed_base, double* out) {
    OrcWorld* W = (OrcWorld*)h;
    if (!W->from_desc || material < 0 || material >= (int32_t)W->mats.size()) return -1;
    MaterialPtr m = material_from_desc(W->foreign, material, W->mats, W->texs);
    if (!m) return -1;
    for (int64_t i = 0; i < n; i++) {
        const double* q = in + 15 * i;
        double* o = out + 11 * i;
        Ray ray{Point3(q[0], q[1], q[2]), Vec3(q[3], q[4], q[5])};
        Hit hit;
        hit.p = Point3(q[6], q[7], q[8]), hit.normal = Vec3(q[9], q[10], q[11]);
        hit.u = q[12], hit.v = q[13], hit.front_face = q[14] != 0.0, hit.material = m.get();
        Ctx cx;
        cx.rng = Pcg64::seed_from_u64(seed_base + (uint64_t)i);
        Pcg64 peek = cx.rng;
        o[10] = peek.unit();
        Color att(0, 0, 0);
        Ray sc{hit.p, Vec3(0, 0, 0)};
        bool alive = m->scatter(ray, hit, cx, att, sc);  // raytrace.rs:92-96: scatter, else emit
        Color em = alive ? Color(0, 0, 0) : m->emit(hit.u, hit.v, hit.p, cx.c);
        o[0] = alive ? 1.0 : 0.0;
        for (int k = 0; k < 3; k++) o[1 + k] = alive ? att.e[k] : 0.0, o[4 + k] = alive ? sc.dir.e[k] : 0.0, o[7 + k] = em.e[k];
    }
    return 0;
}

// ---- small probes used to pin the restatement against known answers ----
void orc_pcg_seed_stream(uint64_t seed, int32_t n, uint64_t* out) {
    Pcg64 r = Pcg64::seed_from_u64(seed);
    for (int i = 0; i < n; i++) out[i] = r.next_u64();
}
void orc_pcg_new_stream(uint64_t state_lo, uint64_t state_hi, uint64_t stream_lo, uint64_t stream_hi, int32_t n,
                        uint64_t* out) {
    Pcg64 r = Pcg64::from_state_stream(((u128)state_hi << 64) | state_lo, ((u128)stream_hi << 64) | stream_lo);
    for (int i = 0; i < n; i++) out[i] = r.next_u64();
}
uint64_t orc_pcg_from_seed_first(const uint8_t seed[32]) {
    Pcg64 r = Pcg64::from_seed(seed);
    return r.next_u64();
}
void orc_pcg_f64_stream(uint64_t seed, double lo, double hi, int32_t n, double* out) {
    Pcg64 r = Pcg64::seed_from_u64(seed);
    for (int i = 0; i < n; i++) out[i] = r.gen_range_f64(lo, hi);
}
uint64_t orc_pcg_usize_stream(uint64_t seed, uint64_t lo, uint64_t hi, int32_t n, uint64_t* out) {
    Pcg64 r = Pcg64::seed_from_u64(seed);
    for (int i = 0; i < n; i++) out[i] = r.gen_range_usize(lo, hi);
    return r.draws;
}
void orc_sphere_uv(const double n[3], double uv[2]) { sphere_uv(Vec3(n[0], n[1], n[2]), uv[0], uv[1]); }
int32_t orc_aabb_hit(const double a[3], const double b[3], const double o[3], const double d[3], double tmin, double tmax) {
    AABB box(Point3(a[0], a[1], a[2]), Point3(b[0], b[1], b[2]));
    return box.hit(Ray{Point3(o[0], o[1], o[2]), Vec3(d[0], d[1], d[2])}, tmin, tmax) ? 1 : 0;
}
void orc_aabb_corners(const double a[3], const double b[3], double mn[3], double mx[3]) {
    AABB box(Point3(a[0], a[1], a[2]), Point3(b[0], b[1], b[2]));
    for (int i = 0; i < 3; i++) {
        mn[i] = box.minimum.e[i];
        mx[i] = box.maximum.e[i];
    }
}
void orc_to_rgb(const double c[3], int32_t spp, int32_t out[3]) { to_rgb(Color(c[0], c[1], c[2]), spp, out); }
void orc_camera_ray(const RtCamera* cam, double s, double t, uint64_t seed, double o[3], double d[3]) {
    Camera c(Point3(cam->lookfrom[0], cam->lookfrom[1], cam->lookfrom[2]),
             Point3(cam->lookat[0], cam->lookat[1], cam->lookat[2]), Vec3(cam->vup[0], cam->vup[1], cam->vup[2]),
             cam->vfov_deg, cam->aspect_ratio, cam->aperture, cam->focus_dist);
    Pcg64 r = Pcg64::seed_from_u64(seed);
    Ray ray = c.get_ray(s, t, r);
    for (int i = 0; i < 3; i++) {
        o[i] = ray.orig.e[i];
        d[i] = ray.dir.e[i];
    }
}
void orc_camera_basis(const RtCamera* cam, double out[19]) {  // origin, llc, horizontal, vertical, u, v, lens_radius
    Camera c(Point3(cam->lookfrom[0], cam->lookfrom[1], cam->lookfrom[2]),
             Point3(cam->lookat[0], cam->lookat[1], cam->lookat[2]), Vec3(cam->vup[0], cam->vup[1], cam->vup[2]),
             cam->vfov_deg, cam->aspect_ratio, cam->aperture, cam->focus_dist);
    const Vec3* vs[6] = {&c.origin, &c.lower_left_corner, &c.horizontal, &c.vertical, &c.u, &c.v};
    for (int i = 0; i < 6; i++)
        for (int k = 0; k < 3; k++) out[3 * i + k] = vs[i]->e[k];
    out[18] = c.lens_radius;
}
// Texture::value of description texture `tex` at (u,v,p)
int32_t orc_texture_value(void* h, int32_t tex, const double* uvp, int64_t n, double* rgb) {
    OrcWorld* W = (OrcWorld*)h;
    if (!W->from_desc || tex < 0 || (size_t)tex >= W->texs.size()) return -1;
    std::shared_ptr<Texture> t = texture_from_desc(W->foreign, tex, W->texs);
    Counters c;
    for (int64_t i = 0; i < n; i++) {
        const double* q = uvp + 5 * i;
        Color col = t->value(q[0], q[1], Point3(q[2], q[3], q[4]), c);
        for (int k = 0; k < 3; k++) rgb[3 * i + k] = col.e[k];
    }
    return 0;
}

}  // extern "C"
