// ORACLE — test infrastructure only.  Nothing under oracle/ is part of the product path.
//
// CPU restatement (f64) of the reference's camera, integrator and render driver:
//   src/camera.rs:15-48, src/raytrace.rs:29-48 (backgrounds), :59-68 (to_rgb), :79-101 (trace_internal),
//   :162-198 (render / render_line / render_pixel), src/rngator.rs:27-31 (per-row Pcg64 streams).
// It stands in for the Rust binary, which cannot be built in this image (no rustc/cargo).
#pragma once
#include <atomic>
#include <thread>

#include "scene.hpp"

namespace orc {

struct Camera {  // camera.rs:3-12
    Point3 origin, lower_left_corner;
    Vec3 horizontal, vertical, u, v;
    double lens_radius;
    double time0 = 0.0, time1 = 0.0;  // EXTENSION (The Next Week's shutter); 0, 0 = the reference's camera, no extra draw

    Camera() : lens_radius(0) {}
    Camera(Point3 lookfrom, Point3 lookat, Vec3 vup, double vfov, double aspect_ratio, double aperture,
           double focus_dist) {  // camera.rs:15-38
        double theta = vfov * PI / 180.0;
        double h = std::tan(theta / 2.0);
        double viewport_height = 2.0 * h;
        double viewport_width = aspect_ratio * viewport_height;
        Vec3 w = (lookfrom - lookat).unit();
        u = vup.cross(w).unit();
        v = w.cross(u);
        origin = lookfrom;
        horizontal = (focus_dist * viewport_width) * u;
        vertical = (focus_dist * viewport_height) * v;
        lower_left_corner = origin - horizontal / 2.0 - vertical / 2.0 - focus_dist * w;
        lens_radius = aperture / 2.0;
    }
    Ray get_ray(double s, double t, Pcg64& rng) const {  // camera.rs:40-48 — the disk sample is always drawn
        Vec3 rd = lens_radius * random_in_unit_disk(rng);
        Vec3 offset = u * rd.x() + v * rd.y();
        double time = time1 > time0 ? rng.gen_range_f64(time0, time1) : time0;
        return Ray{origin + offset, lower_left_corner + s * horizontal + t * vertical - origin - offset, time};
    }
};

struct Background {
    int kind = 0;  // RT_BG_*
    Color top, bottom;
    Color color(const Ray& r) const {  // raytrace.rs:29-35, :44-48
        if (kind == 0) return Color(0, 0, 0);
        Vec3 ud = r.dir.unit();
        double t = 0.5 * (ud.y() + 1.0);
        return (1.0 - t) * bottom + t * top;
    }
};

inline int32_t as_i32(double x) {  // Rust `as i32`: saturating, NaN -> 0
    if (x != x) return 0;
    if (x >= 2147483647.0) return 2147483647;
    if (x <= -2147483648.0) return (int32_t)-2147483647 - 1;
    return (int32_t)x;
}
inline double clamp_f64(double x, double lo, double hi) {  // f64::clamp: NaN stays NaN
    if (x < lo) return lo;
    if (x > hi) return hi;
    return x;
}
inline void to_rgb(const Color& c, int spp, int32_t out[3]) {  // raytrace.rs:59-68
    double scale = 1.0 / (double)spp;
    for (int k = 0; k < 3; k++) {
        double x = std::sqrt(c.e[k] * scale);
        out[k] = as_i32(255.999 * clamp_f64(x, 0.0, 0.99999999));
    }
}

// raytrace.rs:79-101 — recursive, radiance = product of attenuations x one terminal term
inline Color trace(const Ray& ray, const Hittable& world, const Background& bg, int depth, Ctx& cx) {
    if (depth <= 0) {
        cx.c.depth_exhausted++;
        return Color(0, 0, 0);
    }
    cx.c.rays++;
    Hit h;
    if (world.hit(ray, 0.001, INF, cx, h)) {
        Color att;
        Ray scattered;
        cx.c.scatter[h.material->kind]++;
        if (h.material->scatter(ray, h, cx, att, scattered)) {
            return att * trace(scattered, world, bg, depth - 1, cx);
        }
        return h.material->emit(h.u, h.v, h.p, cx.c);
    }
    cx.c.background_evals++;
    return bg.color(ray);
}

struct RenderJob {
    const Hittable* world;
    Background bg;
    Camera cam;
    int width, height, spp, max_depth;
    uint64_t render_seed;  // row j uses Pcg64::seed_from_u64(render_seed + j)  (raytrace.rs:179)
    int row_begin = 0, row_end = -1;  // bounded samples for the CPU baseline: rows [row_begin,row_end)
};

// raytrace.rs:188-198 for one pixel; sum (not yet divided) goes to accum, quantised colour to rgb
inline void render_pixel(const RenderJob& J, int i, int j, Ctx& cx, double sum[3], int32_t rgb[3]) {
    Color pixel(0, 0, 0);
    for (int s = 0; s < J.spp; s++) {
        double u = ((double)i + cx.rng.unit()) / ((double)J.width - 1.0);
        double v = ((double)j + cx.rng.unit()) / ((double)J.height - 1.0);
        Ray r = J.cam.get_ray(u, v, cx.rng);
        cx.c.paths++;
        pixel = pixel + trace(r, *J.world, J.bg, J.max_depth, cx);
    }
    for (int k = 0; k < 3; k++) sum[k] = pixel.e[k];
    to_rgb(pixel, J.spp, rgb);
}

// raytrace.rs:172-186: rows in parallel (dynamic), one RNG per row.  accum/rgb are W*H*3, row j = 0 BOTTOM.
inline void render(const RenderJob& J, int n_threads, double* accum, int32_t* rgb, Counters* counters) {
    int r0 = J.row_begin, r1 = J.row_end < 0 ? J.height : J.row_end;
    std::atomic<int> next(r0);
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads <= 0) n_threads = 1;
    std::vector<Counters> per_thread(n_threads);
    auto worker = [&](int tid) {
        for (;;) {
            int j = next.fetch_add(1);
            if (j >= r1) break;
            Ctx cx;
            cx.rng = Pcg64::seed_from_u64(J.render_seed + (uint64_t)j);
            for (int i = 0; i < J.width; i++) {
                double sum[3];
                int32_t q[3];
                render_pixel(J, i, j, cx, sum, q);
                size_t o = 3 * ((size_t)j * J.width + i);
                for (int k = 0; k < 3; k++) {
                    if (accum) accum[o + k] = sum[k];
                    if (rgb) rgb[o + k] = q[k];
                }
            }
            per_thread[tid].add(cx.c);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; t++) pool.emplace_back(worker, t);
    worker(0);
    for (auto& t : pool) t.join();
    if (counters)
        for (auto& c : per_thread) counters->add(c);
}

}  // namespace orc
