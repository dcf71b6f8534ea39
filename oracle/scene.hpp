// ORACLE — test infrastructure only.  Nothing under oracle/ is part of the product path.
//
// CPU restatement (C++17, f64) of the reference's geometry, acceleration structure, materials and
// textures.  Every class cites the reference file:line it follows; quirks are kept on purpose
// (SURVEY App. C).  Compile with -ffp-contract=off: Rust never fuses a*b+c.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <vector>

#include "pcg64.hpp"

namespace orc {

static const double PI = 3.14159265358979323846264338327950288;  // std::f64::consts::PI
static const double INF = std::numeric_limits<double>::infinity();

// ---- src/vec.rs ----
struct Vec3 {
    double e[3];
    Vec3() : e{0, 0, 0} {}
    Vec3(double a, double b, double c) : e{a, b, c} {}
    double x() const { return e[0]; }
    double y() const { return e[1]; }
    double z() const { return e[2]; }
    double dot(const Vec3& v) const { return e[0] * v.e[0] + e[1] * v.e[1] + e[2] * v.e[2]; }  // vec.rs:71-73
    double length_squared() const { return e[0] * e[0] + e[1] * e[1] + e[2] * e[2]; }          // vec.rs:83-85
    double length() const { return std::sqrt(length_squared()); }                              // vec.rs:63-65
    Vec3 cross(const Vec3& v) const {                                                          // vec.rs:75-81
        return Vec3(e[1] * v.e[2] - e[2] * v.e[1], e[2] * v.e[0] - e[0] * v.e[2], e[0] * v.e[1] - e[1] * v.e[0]);
    }
    Vec3 unit() const {  // vec.rs:67-69: divide each component by the length
        double l = length();
        return Vec3(e[0] / l, e[1] / l, e[2] / l);
    }
    bool near_zero() const {  // vec.rs:54-57
        const double S = 1e-8;
        return std::fabs(e[0]) < S && std::fabs(e[1]) < S && std::fabs(e[2]) < S;
    }
};
inline Vec3 operator-(const Vec3& a) { return Vec3(-a.e[0], -a.e[1], -a.e[2]); }
inline Vec3 operator+(const Vec3& a, const Vec3& b) { return Vec3(a.e[0] + b.e[0], a.e[1] + b.e[1], a.e[2] + b.e[2]); }
inline Vec3 operator-(const Vec3& a, const Vec3& b) { return Vec3(a.e[0] - b.e[0], a.e[1] - b.e[1], a.e[2] - b.e[2]); }
inline Vec3 operator*(const Vec3& a, double s) { return Vec3(a.e[0] * s, a.e[1] * s, a.e[2] * s); }
inline Vec3 operator*(double s, const Vec3& a) { return a * s; }
inline Vec3 operator*(const Vec3& a, const Vec3& b) { return Vec3(a.e[0] * b.e[0], a.e[1] * b.e[1], a.e[2] * b.e[2]); }
inline Vec3 operator/(const Vec3& a, double s) { return Vec3(a.e[0] / s, a.e[1] / s, a.e[2] / s); }
typedef Vec3 Point3;
typedef Vec3 Color;

struct Ray {  // vec.rs:215-228 — direction NOT normalised.  The reference's Ray has no time; `time` exists only for the
              // moving-sphere EXTENSION below (it stays 0 for everything the reference can express)
    Point3 orig;
    Vec3 dir;
    double time = 0.0;
    Point3 at(double t) const { return orig + t * dir; }
};

// instrumentation: how much work the REFERENCE's algorithm does.  Not part of the algorithm.
struct Counters {
    uint64_t paths = 0, rays = 0;
    uint64_t aabb_tests = 0, sphere_tests = 0, rect_tests = 0, xform = 0, medium_tests = 0;
    uint64_t scatter[6] = {0, 0, 0, 0, 0, 0};  // by RT_MAT_* kind (index 0 unused)
    uint64_t perlin_evals = 0, image_evals = 0, background_evals = 0, depth_exhausted = 0;
    void add(const Counters& o) {
        paths += o.paths; rays += o.rays; aabb_tests += o.aabb_tests; sphere_tests += o.sphere_tests;
        rect_tests += o.rect_tests; xform += o.xform; medium_tests += o.medium_tests;
        for (int i = 0; i < 6; i++) scatter[i] += o.scatter[i];
        perlin_evals += o.perlin_evals; image_evals += o.image_evals; background_evals += o.background_evals;
        depth_exhausted += o.depth_exhausted;
    }
};

// what the reference threads through every call as `rng: &mut dyn RngCore`
struct Ctx {
    Pcg64 rng;
    Counters c;
    bool skip_media = false;  // parity harness only: surfaces-only closest hit (media are stochastic)
};

// vec.rs:15-52 samplers
inline Vec3 random_vec(double lo, double hi, Pcg64& r) {
    double a = r.gen_range_f64(lo, hi);
    double b = r.gen_range_f64(lo, hi);
    double c = r.gen_range_f64(lo, hi);
    return Vec3(a, b, c);
}
inline Vec3 random_in_unit_sphere(Pcg64& r) {  // vec.rs:23-30
    for (;;) {
        Vec3 p = random_vec(-1.0, 1.0, r);
        if (p.length_squared() < 1.0) return p;
    }
}
inline Vec3 random_in_hemisphere(const Vec3& n, Pcg64& r) {  // vec.rs:36-43
    Vec3 p = random_in_unit_sphere(r);
    return n.dot(p) > 0.0 ? p : -p;
}
inline Vec3 random_in_unit_disk(Pcg64& r) {  // vec.rs:45-52
    for (;;) {
        double a = r.gen_range_f64(-1.0, 1.0);
        double b = r.gen_range_f64(-1.0, 1.0);
        Vec3 p(a, b, 0.0);
        if (p.length_squared() < 1.0) return p;
    }
}

// ---- src/textures.rs, src/image_texture.rs ----
struct Texture {
    virtual ~Texture() {}
    virtual Color value(double u, double v, const Point3& p, Counters& c) const = 0;
};
struct SolidColor : Texture {  // textures.rs:8-26
    Color color;
    explicit SolidColor(Color c) : color(c) {}
    Color value(double, double, const Point3&, Counters&) const override { return color; }
};
struct Checker : Texture {  // textures.rs:28-49
    std::shared_ptr<Texture> odd, even;
    Checker(std::shared_ptr<Texture> o, std::shared_ptr<Texture> e) : odd(o), even(e) {}
    Color value(double u, double v, const Point3& p, Counters& c) const override {
        double sines = std::sin(5.0 * p.x()) * std::sin(5.0 * p.y()) * std::sin(5.0 * p.z());
        return sines < 0.0 ? odd->value(u, v, p, c) : even->value(u, v, p, c);
    }
};
struct Perlin {  // textures.rs:51-149 — 1024 points, not the book's 256
    static const int N = 1024;
    Vec3 ranvec[N];
    int64_t perm_x[N], perm_y[N], perm_z[N];
    int turbulence_depth = 7;

    explicit Perlin(Pcg64& rng) {  // textures.rs:62-74
        for (int i = 0; i < N; i++) ranvec[i] = random_vec(-1.0, 1.0, rng).unit();
        permute(rng, perm_x);
        permute(rng, perm_y);
        permute(rng, perm_z);
    }
    Perlin() {}
    static void permute(Pcg64& rng, int64_t* out) {  // textures.rs:136-148 — gen_range(0..i), i exclusive
        for (int i = 0; i < N; i++) out[i] = i;
        for (int i = N - 1; i >= 1; i--) {
            uint64_t j = rng.gen_range_usize(0, (uint64_t)i);
            std::swap(out[i], out[j]);
        }
    }
    static int64_t rem_euclid(int64_t a, int64_t m) {
        int64_t r = a % m;
        return r < 0 ? r + m : r;
    }
    double noise(const Point3& p) const {  // textures.rs:90-113
        double fx = std::floor(p.x()), fy = std::floor(p.y()), fz = std::floor(p.z());
        double u = p.x() - fx, v = p.y() - fy, w = p.z() - fz;
        int64_t i = (int64_t)fx, j = (int64_t)fy, k = (int64_t)fz;
        Vec3 c[2][2][2];
        for (int di = 0; di < 2; di++)
            for (int dj = 0; dj < 2; dj++)
                for (int dk = 0; dk < 2; dk++) {
                    int64_t ii = rem_euclid(i + di, N), jj = rem_euclid(j + dj, N), kk = rem_euclid(k + dk, N);
                    c[di][dj][dk] = ranvec[perm_x[ii] ^ perm_y[jj] ^ perm_z[kk]];
                }
        return trilinear(c, u, v, w);
    }
    static double trilinear(const Vec3 c[2][2][2], double u, double v, double w) {  // textures.rs:115-134
        double uu = u * u * (3.0 - 2.0 * u);
        double vv = v * v * (3.0 - 2.0 * v);
        double ww = w * w * (3.0 - 2.0 * w);
        double accum = 0.0;
        for (int i = 0; i < 2; i++)
            for (int j = 0; j < 2; j++)
                for (int k = 0; k < 2; k++) {
                    Vec3 weight(u - i, v - j, w - k);
                    accum += ((double)i * uu + (double)(1 - i) * (1.0 - uu)) * ((double)j * vv + (double)(1 - j) * (1.0 - vv)) *
                             ((double)k * ww + (double)(1 - k) * (1.0 - ww)) * weight.dot(c[i][j][k]);
                }
        return accum;
    }
    double turbulence(const Point3& p) const {  // textures.rs:76-88
        double accum = 0.0, weight = 1.0;
        Point3 tp = p;
        for (int i = 0; i < turbulence_depth; i++) {
            accum += weight * noise(tp);
            weight *= 0.5;
            tp = 2.0 * tp;
        }
        return std::fabs(accum);
    }
};
struct NoiseTexture : Texture {  // textures.rs:151-167 — marble phase on z
    std::shared_ptr<Perlin> noise;
    double scale;
    NoiseTexture(std::shared_ptr<Perlin> n, double s) : noise(n), scale(s) {}
    Color value(double, double, const Point3& p, Counters& c) const override {
        c.perlin_evals++;
        // Color::ONE * 0.5 * (1 + sin(..)) evaluates left to right: (1*0.5) then one multiply per channel
        return (Color(1.0, 1.0, 1.0) * 0.5) * (1.0 + std::sin(scale * p.z() + 10.0 * noise->turbulence(scale * p)));
    }
};
struct ImageTexture : Texture {  // image_texture.rs:16-28 — nearest texel, v flipped
    int width, height;
    std::shared_ptr<std::vector<uint8_t>> rgb;
    ImageTexture(int w, int h, std::shared_ptr<std::vector<uint8_t>> d) : width(w), height(h), rgb(d) {}
    static double clamp01(double x) { return x < 0.0 ? 0.0 : (x > 1.0 ? 1.0 : x); }  // f64::clamp keeps NaN
    static uint32_t sat_u32(double x) {  // Rust `as u32`: saturating, NaN -> 0
        if (!(x > 0.0)) return 0;
        if (x >= 4294967295.0) return 4294967295u;
        return (uint32_t)x;
    }
    Color value(double u, double v, const Point3&, Counters& c) const override {
        c.image_evals++;
        u = clamp01(u);
        v = clamp01(1.0 - v);
        uint32_t i = sat_u32(u * (double)width), j = sat_u32(v * (double)height);
        if (i > (uint32_t)width - 1) i = width - 1;
        if (j > (uint32_t)height - 1) j = height - 1;
        const uint8_t* px = rgb->data() + 3 * ((size_t)j * width + i);
        return Color(px[0] / 255.0, px[1] / 255.0, px[2] / 255.0);
    }
};

// ---- src/hittable.rs ----
struct Material;
struct Hit {  // hittable.rs:7-15
    Point3 p;
    Vec3 normal;
    double t = 0, u = 0, v = 0;
    bool front_face = false;
    const Material* material = nullptr;
    int32_t desc_node = -1;  // instrumentation: description node of the primitive that was hit
};
inline Hit face_normal_hit(const Point3& p, double t, double u, double v, const Vec3& outward, const Ray& r,
                           const Material* m) {  // hittable.rs:18-30
    Hit h;
    h.front_face = outward.dot(r.dir) < 0.0;
    h.normal = h.front_face ? outward : -outward;
    h.p = p; h.t = t; h.u = u; h.v = v; h.material = m;
    return h;
}

// ---- src/materials.rs, src/volumes.rs:67-83 ----
struct Material {
    int kind = 0;        // RT_MAT_*
    int desc_index = -1;  // index in the scene description (instrumentation / hashing)
    virtual ~Material() {}
    virtual bool scatter(const Ray& ray, const Hit& h, Ctx& cx, Color& attenuation, Ray& scattered) const = 0;
    virtual Color emit(double, double, const Point3&, Counters&) const { return Color(0, 0, 0); }  // materials.rs:9-11
};
struct Lambertian : Material {  // materials.rs:14-34
    std::shared_ptr<Texture> albedo;
    explicit Lambertian(std::shared_ptr<Texture> a) : albedo(a) { kind = 1; }
    bool scatter(const Ray& ray, const Hit& h, Ctx& cx, Color& att, Ray& out) const override {
        Vec3 dir = h.normal + random_in_hemisphere(h.normal, cx.rng);
        if (dir.near_zero()) dir = h.normal;
        att = albedo->value(h.u, h.v, h.p, cx.c);
        out = Ray{h.p, dir, ray.time};
        return true;
    }
};
inline Vec3 reflect(const Vec3& v, const Vec3& n) { return v - 2.0 * v.dot(n) * n; }  // materials.rs:47-49
struct Metal : Material {  // materials.rs:36-61
    Color albedo;
    double fuzz;
    Metal(Color a, double f) : albedo(a), fuzz(f) { kind = 2; }
    bool scatter(const Ray& ray, const Hit& h, Ctx& cx, Color& att, Ray& out) const override {
        Vec3 reflected = reflect(ray.dir.unit(), h.normal);
        out = Ray{h.p, reflected + fuzz * random_in_unit_sphere(cx.rng), ray.time};
        if (out.dir.dot(h.normal) > 0.0) {
            att = albedo;
            return true;
        }
        return false;
    }
};
inline Vec3 refract(const Vec3& uv, const Vec3& n, double etai_over_etat) {  // materials.rs:63-68
    double cos_theta = std::fmin((-uv).dot(n), 1.0);
    Vec3 r_out_perp = etai_over_etat * (uv + cos_theta * n);
    Vec3 r_out_parallel = -std::sqrt(std::fabs(1.0 - r_out_perp.length_squared())) * n;
    return r_out_perp + r_out_parallel;
}
inline double powi5(double x) {  // f64::powi(5): square-and-multiply
    double x2 = x * x;
    return x * (x2 * x2);
}
inline double reflectance(double cos_theta, double ratio) {  // materials.rs:81-86 (Schlick)
    double r0 = (1.0 - ratio) / (1.0 + ratio);
    double r0_sq = r0 * r0;
    return r0_sq + (1.0 - r0_sq) * powi5(1.0 - cos_theta);
}
struct Dielectric : Material {  // materials.rs:70-106
    double ior;
    explicit Dielectric(double i) : ior(i) { kind = 3; }
    bool scatter(const Ray& ray, const Hit& h, Ctx& cx, Color& att, Ray& out) const override {
        att = Color(1.0, 1.0, 1.0);
        double ratio = !h.front_face ? ior : 1.0 / ior;
        Vec3 ud = ray.dir.unit();
        double cos_theta = std::fmin(h.normal.dot(-ud), 1.0);
        double sin_theta = std::sqrt(1.0 - cos_theta * cos_theta);
        bool cannot_refract = ratio * sin_theta > 1.0;
        // `||` short-circuits: no draw on total internal reflection (materials.rs:98)
        Vec3 dir = (cannot_refract || reflectance(cos_theta, ratio) > cx.rng.unit()) ? reflect(ud, h.normal)
                                                                                      : refract(ud, h.normal, ratio);
        out = Ray{h.p, dir, ray.time};
        return true;
    }
};
struct DiffuseLight : Material {  // materials.rs:108-127 — emits from both faces
    std::shared_ptr<Texture> texture;
    explicit DiffuseLight(std::shared_ptr<Texture> t) : texture(t) { kind = 4; }
    bool scatter(const Ray&, const Hit&, Ctx&, Color&, Ray&) const override { return false; }
    Color emit(double u, double v, const Point3& p, Counters& c) const override { return texture->value(u, v, p, c); }
};
struct Isotropic : Material {  // volumes.rs:67-83 — direction = raw in-ball point
    std::shared_ptr<Texture> albedo;
    explicit Isotropic(std::shared_ptr<Texture> a) : albedo(a) { kind = 5; }
    bool scatter(const Ray& ray, const Hit& h, Ctx& cx, Color& att, Ray& out) const override {
        out = Ray{h.p, random_in_unit_sphere(cx.rng), ray.time};
        att = albedo->value(h.u, h.v, h.p, cx.c);
        return true;
    }
};

// ---- src/bhv.rs:8-53 ----
struct AABB {
    Point3 minimum, maximum;
    AABB() {}
    AABB(const Point3& a, const Point3& b) {  // bhv.rs:16-20 (f64::min/max ignore NaN like fmin/fmax)
        for (int i = 0; i < 3; i++) {
            minimum.e[i] = std::fmin(a.e[i], b.e[i]);
            maximum.e[i] = std::fmax(a.e[i], b.e[i]);
        }
    }
    bool hit(const Ray& r, double tmin, double tmax) const {  // bhv.rs:29-42 — divides, not reciprocals
        for (int a = 0; a < 3; a++) {
            double t0 = (minimum.e[a] - r.orig.e[a]) / r.dir.e[a];
            double t1 = (maximum.e[a] - r.orig.e[a]) / r.dir.e[a];
            tmin = std::fmax(std::fmin(t0, t1), tmin);
            tmax = std::fmin(std::fmax(t0, t1), tmax);
            if (tmax <= tmin) return false;
        }
        return true;
    }
    AABB surround(const AABB& o) const {  // bhv.rs:44-52
        Point3 mn, mx;
        for (int a = 0; a < 3; a++) {
            mn.e[a] = std::fmin(minimum.e[a], o.minimum.e[a]);
            mx.e[a] = std::fmax(maximum.e[a], o.maximum.e[a]);
        }
        return AABB(mn, mx);
    }
};

struct Hittable {  // hittable.rs:33-35
    int32_t desc_node = -1;
    virtual ~Hittable() {}
    virtual bool hit(const Ray& r, double t_min, double t_max, Ctx& cx, Hit& out) const = 0;
    virtual AABB bounding_box() const = 0;  // bhv.rs:61-63 (Bounded); unbounded kinds return an empty box
};
typedef std::shared_ptr<Hittable> HittablePtr;
typedef std::shared_ptr<Material> MaterialPtr;

struct Empty : Hittable {  // shapes.rs:8-24
    bool hit(const Ray&, double, double, Ctx&, Hit&) const override { return false; }
    AABB bounding_box() const override { return AABB(Point3(), Point3()); }
};

struct HittableList : Hittable {  // hittable.rs:37-68
    std::vector<HittablePtr> contents;
    bool hit(const Ray& r, double t_min, double t_max, Ctx& cx, Hit& out) const override {
        bool any = false;
        double closest = t_max;
        Hit h;
        for (const auto& o : contents) {
            if (o->hit(r, t_min, closest, cx, h)) {
                closest = h.t;
                out = h;
                any = true;
            }
        }
        return any;
    }
    AABB bounding_box() const override { return AABB(); }
};

// ---- src/shapes.rs ----
inline void sphere_uv(const Vec3& n, double& u, double& v) {  // shapes.rs:44-55
    double theta = std::acos(-n.y());
    double phi = std::atan2(-n.z(), n.x()) + PI;
    u = phi / (2.0 * PI);
    v = theta / PI;
}
struct Sphere : Hittable {  // shapes.rs:26-90
    Point3 center;
    double radius;
    MaterialPtr material;
    Sphere(Point3 c, double r, MaterialPtr m) : center(c), radius(r), material(m) {}
    bool hit(const Ray& r, double t_min, double t_max, Ctx& cx, Hit& out) const override {
        cx.c.sphere_tests++;
        Vec3 oc = r.orig - center;
        double a = r.dir.length_squared();
        double half_b = oc.dot(r.dir);
        double c = oc.length_squared() - radius * radius;
        double disc = half_b * half_b - a * c;
        if (disc < 0.0) return false;
        double sqrtd = std::sqrt(disc);
        double root = (-half_b - sqrtd) / a;
        if (root < t_min || t_max < root) {
            root = (-half_b + sqrtd) / a;
            if (root < t_min || t_max < root) return false;
        }
        Point3 p = r.at(root);
        Vec3 normal = (p - center) / radius;  // signed radius flips the normal (negative-radius shells)
        double u, v;
        sphere_uv(normal, u, v);  // runs on every sphere hit, needed or not
        out = face_normal_hit(p, root, u, v, normal, r, material.get());
        out.desc_node = desc_node;
        return true;
    }
    AABB bounding_box() const override {  // shapes.rs:84-89
        Vec3 rv(radius, radius, radius);
        return AABB(center - rv, center + rv);
    }
};

// ---- EXTENSION, not in the reference: the MovingSphere of "Ray Tracing: The Next Week" (section 2.2-2.5), restated so that
// the device's RT_NODE_MOVING_SPHERE has something to be compared with.  centre(time) = c0 + time * (c1 - c0), time in
// [0, 1]; the hit is Sphere::hit with that centre; the box surrounds both ends.
struct MovingSphere : Hittable {
    Point3 center0, center1;
    double radius;
    MaterialPtr material;
    MovingSphere(Point3 c0, Point3 c1, double r, MaterialPtr m) : center0(c0), center1(c1), radius(r), material(m) {}
    bool hit(const Ray& r, double t_min, double t_max, Ctx& cx, Hit& out) const override {
        Sphere at_time(center0 + r.time * (center1 - center0), radius, material);
        at_time.desc_node = desc_node;
        return at_time.hit(r, t_min, t_max, cx, out);
    }
    AABB bounding_box() const override {
        Vec3 rv(std::fabs(radius), std::fabs(radius), std::fabs(radius));
        return AABB(center0 - rv, center0 + rv).surround(AABB(center1 - rv, center1 + rv));
    }
};

// ---- src/aarects.rs ----
struct AARect {
    int a0, a1, aplane;
    double a0_v0, a0_v1, a1_v0, a1_v1, aplane_v;
    AARect() {}
    AARect(int ax0, double v00, double v01, int ax1, double v10, double v11, double k) {  // aarects.rs:31-43
        a0 = ax0; a1 = ax1;
        aplane = 3 - ax0 - ax1;  // other(a0,a1)
        a0_v0 = std::fmin(v00, v01);
        a0_v1 = std::fmax(v01, v00);
        a1_v0 = std::fmin(v10, v10);  // sic: a1_v0.min(a1_v0)  (aarects.rs:39)
        a1_v1 = std::fmax(v11, v10);
        aplane_v = k;
    }
    bool hit(const Ray& r, double tmin, double tmax, const Material* m, Hit& out) const {  // aarects.rs:45-64
        double t = (aplane_v - r.orig.e[aplane]) / r.dir.e[aplane];
        if (t < tmin || t > tmax) return false;
        double a0_v = r.orig.e[a0] + t * r.dir.e[a0];
        double a1_v = r.orig.e[a1] + t * r.dir.e[a1];
        if (a0_v < a0_v0 || a0_v > a0_v1 || a1_v < a1_v0 || a1_v > a1_v1) return false;
        double u = (a0_v - a0_v0) / (a0_v1 - a0_v0);
        double v = (a1_v - a1_v0) / (a1_v1 - a1_v0);
        Vec3 n;
        n.e[aplane] = 1.0;
        out = face_normal_hit(r.at(t), t, u, v, n, r, m);
        return true;
    }
    AABB bounding_box() const {  // aarects.rs:66-77 — sic: maximum uses the *_v0 bounds (degenerate box)
        Point3 mn, mx;
        mn.e[a0] = a0_v0; mn.e[a1] = a1_v0; mn.e[aplane] = aplane_v - 0.001;
        mx.e[a0] = a0_v0; mx.e[a1] = a1_v0; mx.e[aplane] = aplane_v + 0.001;
        return AABB(mn, mx);
    }
};
struct RectShape : Hittable {  // shapes.rs:92-164  XYRect / XZRect / YZRect
    AARect r;
    MaterialPtr material;
    RectShape(AARect rr, MaterialPtr m) : r(rr), material(m) {}
    bool hit(const Ray& ray, double tmin, double tmax, Ctx& cx, Hit& out) const override {
        cx.c.rect_tests++;
        if (!r.hit(ray, tmin, tmax, material.get(), out)) return false;
        out.desc_node = desc_node;
        return true;
    }
    AABB bounding_box() const override { return r.bounding_box(); }
};
inline AARect xy_rect(double x0, double x1, double y0, double y1, double z) { return AARect(0, x0, x1, 1, y0, y1, z); }
inline AARect xz_rect(double x0, double x1, double z0, double z1, double y) { return AARect(0, x0, x1, 2, z0, z1, y); }
inline AARect yz_rect(double y0, double y1, double z0, double z1, double x) { return AARect(1, y0, y1, 2, z0, z1, x); }

struct Block : Hittable {  // shapes.rs:166-198 — six rects in a list, own (min,max) box
    Point3 mn, mx;
    AARect sides[6];
    MaterialPtr material;
    Block(Point3 p0, Point3 p1, MaterialPtr m) : mn(p0), mx(p1), material(m) {
        sides[0] = xy_rect(p0.x(), p1.x(), p0.y(), p1.y(), p1.z());
        sides[1] = xy_rect(p0.x(), p1.x(), p0.y(), p1.y(), p0.z());
        sides[2] = xz_rect(p0.x(), p1.x(), p0.z(), p1.z(), p0.y());
        sides[3] = xz_rect(p0.x(), p1.x(), p0.z(), p1.z(), p1.y());
        sides[4] = yz_rect(p0.y(), p1.y(), p0.z(), p1.z(), p0.x());
        sides[5] = yz_rect(p0.y(), p1.y(), p0.z(), p1.z(), p1.x());
    }
    bool hit(const Ray& r, double tmin, double tmax, Ctx& cx, Hit& out) const override {
        bool any = false;
        double closest = tmax;
        Hit h;
        for (int i = 0; i < 6; i++) {  // HittableList::hit over the six sides
            cx.c.rect_tests++;
            if (sides[i].hit(r, tmin, closest, material.get(), h)) {
                closest = h.t;
                out = h;
                any = true;
            }
        }
        if (any) out.desc_node = desc_node;
        return any;
    }
    AABB bounding_box() const override { return AABB(mn, mx); }
};

// ---- src/transforms.rs ----
struct Translate : Hittable {  // transforms.rs:20-49
    HittablePtr original;
    Vec3 offset;
    Translate(Vec3 o, HittablePtr h) : original(h), offset(o) {}
    bool hit(const Ray& r, double t_min, double t_max, Ctx& cx, Hit& out) const override {
        cx.c.xform++;
        Ray moved{r.orig - offset, r.dir, r.time};
        Hit h;
        if (!original->hit(moved, t_min, t_max, cx, h)) return false;
        // face-forwarding re-applied to an already flipped normal: front_face ends up true (App. C 16)
        int32_t dn = h.desc_node;
        out = face_normal_hit(h.p + offset, h.t, h.u, h.v, h.normal, moved, h.material);
        out.desc_node = dn;
        return true;
    }
    AABB bounding_box() const override {
        AABB b = original->bounding_box();
        return AABB(b.minimum + offset, b.maximum + offset);
    }
};
struct Rotate : Hittable {  // transforms.rs:51-148
    int a1;
    double sin_theta, cos_theta;
    AABB bbox;
    HittablePtr original;
    Rotate(int axis, double angle, HittablePtr h) : original(h) {  // transforms.rs:59-97
        a1 = axis;
        int a2 = (a1 + 1) % 3, a0 = (a1 + 2) % 3;
        double theta = angle * PI / 180.0;
        sin_theta = std::sin(theta);
        cos_theta = std::cos(theta);
        AABB b = original->bounding_box();
        Point3 mn(-INF, -INF, -INF), mx(-INF, -INF, -INF);  // sic: min also starts at -inf (App. C 14)
        for (int i = 0; i < 2; i++)
            for (int j = 0; j < 2; j++)
                for (int k = 0; k < 2; k++) {
                    double a0_v = i == 1 ? b.maximum.e[a0] : b.minimum.e[a0];
                    double a1_v = j == 1 ? b.maximum.e[a1] : b.minimum.e[a1];
                    double a2_v = k == 1 ? b.maximum.e[a2] : b.minimum.e[a2];
                    double n0 = cos_theta * a0_v + sin_theta * a2_v;
                    double n2 = -sin_theta * a0_v + cos_theta * a2_v;
                    double tester[3];
                    tester[a0] = n0; tester[a1] = a1_v; tester[a2] = n2;
                    for (int c = 0; c < 3; c++) {
                        mn.e[c] = std::fmin(mn.e[c], tester[c]);
                        mx.e[c] = std::fmax(mx.e[c], tester[c]);
                    }
                }
        bbox = AABB(mn, mx);
    }
    Vec3 rotate_back(const Vec3& v) const {  // transforms.rs:106-114
        int a0 = (a1 + 2) % 3, a2 = (a1 + 1) % 3;
        Vec3 r = v;
        r.e[a0] = cos_theta * v.e[a0] - sin_theta * v.e[a2];
        r.e[a2] = sin_theta * v.e[a0] + cos_theta * v.e[a2];
        return r;
    }
    Vec3 rotate(const Vec3& v) const {  // transforms.rs:116-124
        int a0 = (a1 + 2) % 3, a2 = (a1 + 1) % 3;
        Vec3 r = v;
        r.e[a0] = cos_theta * v.e[a0] + sin_theta * v.e[a2];
        r.e[a2] = -sin_theta * v.e[a0] + cos_theta * v.e[a2];
        return r;
    }
    bool hit(const Ray& r, double t_min, double t_max, Ctx& cx, Hit& out) const override {  // transforms.rs:127-142
        cx.c.xform++;
        Ray rr{rotate_back(r.orig), rotate_back(r.dir), r.time};
        Hit h;
        if (!original->hit(rr, t_min, t_max, cx, h)) return false;
        int32_t dn = h.desc_node;
        out = face_normal_hit(rotate(h.p), h.t, h.u, h.v, rotate(h.normal), rr, h.material);
        out.desc_node = dn;
        return true;
    }
    AABB bounding_box() const override { return bbox; }
};

// ---- src/volumes.rs:7-65 ----
struct ConstantMedium : Hittable {
    HittablePtr boundary;
    MaterialPtr phase;  // Isotropic
    double neg_inv_density;
    ConstantMedium(HittablePtr b, double d, MaterialPtr iso) : boundary(b), phase(iso), neg_inv_density(-1.0 / d) {}
    // the clipped boundary interval (deterministic part of hit); exposed for the parity harness
    bool interval(const Ray& r, double t_min, double t_max, Ctx& cx, double& t1, double& t2) const {
        Hit h1, h2;
        if (!boundary->hit(r, -INF, INF, cx, h1)) return false;
        if (!boundary->hit(r, h1.t + 0.001, INF, cx, h2)) return false;
        t1 = std::fmax(h1.t, t_min);
        t2 = std::fmin(h2.t, t_max);
        if (t1 >= t2) return false;
        t1 = std::fmax(t1, 0.0);
        return true;
    }
    bool hit(const Ray& r, double t_min, double t_max, Ctx& cx, Hit& out) const override {
        if (cx.skip_media) return false;
        cx.c.medium_tests++;
        double t1, t2;
        if (!interval(r, t_min, t_max, cx, t1, t2)) return false;
        double ray_scale = r.dir.length();
        double distance_inside = (t2 - t1) * ray_scale;
        double hit_distance = neg_inv_density * std::log(cx.rng.unit());
        if (hit_distance > distance_inside) return false;
        double t = t1 + hit_distance / ray_scale;
        out = Hit();
        out.p = r.at(t);
        out.t = t;
        out.u = 0.0; out.v = 0.0;
        out.normal = Vec3(1.0, 0.0, 0.0);
        out.front_face = true;
        out.material = phase.get();
        out.desc_node = desc_node;
        return true;
    }
    AABB bounding_box() const override { return AABB(); }
};

// ---- src/bhv.rs:84-166 ----
struct BvhNode {
    bool leaf = true;
    HittablePtr shape;  // Leaf
    AABB bounds;        // Inner
    std::unique_ptr<BvhNode> left, right;

    AABB bounding_box() const { return leaf ? shape->bounding_box() : bounds; }

    // Node::new: one usize draw per inner node (pre-order), STABLE sort on box minimum, split len/2
    static std::unique_ptr<BvhNode> build(std::vector<HittablePtr>& shapes, size_t lo, size_t hi, Pcg64& rng,
                                          std::vector<int>* axes_out) {
        std::unique_ptr<BvhNode> n(new BvhNode());
        size_t len = hi - lo;
        if (len == 0) {
            n->shape = std::make_shared<Empty>();
        } else if (len == 1) {
            n->shape = shapes[lo];
        } else {
            int axis = (int)rng.gen_range_usize(0, 3);
            if (axes_out) axes_out->push_back(axis);
            std::stable_sort(shapes.begin() + lo, shapes.begin() + hi, [axis](const HittablePtr& a, const HittablePtr& b) {
                return a->bounding_box().minimum.e[axis] < b->bounding_box().minimum.e[axis];
            });
            size_t mid = lo + len / 2;
            n->leaf = false;
            n->left = build(shapes, lo, mid, rng, axes_out);
            n->right = build(shapes, mid, hi, rng, axes_out);
            n->bounds = n->left->bounding_box().surround(n->right->bounding_box());
        }
        return n;
    }
    // Node::hit: box test on inner nodes only; LEFT always first, RIGHT with t_max = left.t; leaves untested
    bool hit(const Ray& r, double tmin, double tmax, Ctx& cx, Hit& out) const {
        if (leaf) return shape->hit(r, tmin, tmax, cx, out);
        cx.c.aabb_tests++;
        if (!bounds.hit(r, tmin, tmax)) return false;
        Hit hl;
        bool got_left = left->hit(r, tmin, tmax, cx, hl);
        double tmax_right = got_left ? hl.t : tmax;
        Hit hr;
        if (right->hit(r, tmin, tmax_right, cx, hr)) {
            out = hr;
            return true;
        }
        if (got_left) out = hl;
        return got_left;
    }
};
struct BHV : Hittable {
    std::unique_ptr<BvhNode> root;
    std::vector<int> axes;  // instrumentation: the axis drawn at each inner node, pre-order
    BHV(std::vector<HittablePtr> shapes, Pcg64& rng) { root = BvhNode::build(shapes, 0, shapes.size(), rng, &axes); }
    bool hit(const Ray& r, double tmin, double tmax, Ctx& cx, Hit& out) const override {
        return root->hit(r, tmin, tmax, cx, out);
    }
    AABB bounding_box() const override { return root->bounding_box(); }
};

}  // namespace orc
