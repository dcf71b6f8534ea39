// TEST INFRASTRUCTURE ONLY — not part of librt_b200.so and never reachable from the product API.
//
// Compiles the device header (rt_device.cuh) as plain C++ (RTB_HOST_EMULATION) so that the `-m "not gpu"`
// test-suite can run the exact device functions (hit_sphere, hit_box, closest_hit, sample_media, scatter,
// Philox, camera) on the CPU of the build container, which has no GPU, and compare them with the oracle.
// The GPU parity tests (-m gpu) go through the real C ABI instead.
#define RTB_HOST_EMULATION 1
#include <atomic>
#include <cstdio>
#include <string>
#include <thread>
#include <vector>

#include "../../mu-lambda-raytracer_b200/csrc/flatten.h"
#include "../../mu-lambda-raytracer_b200/csrc/rt_device.cuh"

using namespace rtb;

struct EmulScene {
    FlatScene flat;
    std::vector<DImage> images;
    DSceneView view;
};

static void make_view(EmulScene& e) {
    FlatScene& f = e.flat;
    e.images.clear();
    for (auto& im : f.images) e.images.push_back(DImage{(unsigned long long)(uintptr_t)im.rgba.data(), im.width, im.height});
    DSceneView& v = e.view;
    v.nodes = f.nodes.data(), v.nodes4 = f.nodes4.empty() ? nullptr : f.nodes4.data(), v.n_nodes4 = (int)f.nodes4.size(), v.prims = f.prims.data(), v.big = f.big.data(), v.inst = f.inst.data();
    v.mats = f.mats.data(), v.texs = f.texs.data(), v.media = f.media.data(), v.media_prims = f.media_prims.data();
    v.perlin_vec = f.perlin_vec.data(), v.perlin_perm = f.perlin_perm.data(), v.images = e.images.data(), v.moving = f.moving.data();
    v.n_nodes = (int)f.nodes.size(), v.n_prims = (int)f.prims.size(), v.n_media = (int)f.media.size();
    v.n_perlin = (int)(f.perlin_vec.size() / (4 * RTB_PERLIN_POINTS));
    v.media_general = f.media.size() > 4 ? 1 : 0;
    for (const auto& m : f.media) v.media_general |= m.count > 1 ? 1 : 0;
    v.clear_media = f.clear_media;
    v.bg_kind = f.bg_kind;
    for (int k = 0; k < 3; ++k) v.bg_top[k] = f.bg_top[k], v.bg_bottom[k] = f.bg_bottom[k];
}

static thread_local std::string g_err;

extern "C" {

const char* emul_last_error() { return g_err.c_str(); }

void* emul_scene_create(const RtSceneDesc* d, int32_t root, int32_t build_bvh) {
    EmulScene* e = new EmulScene();
    std::string err;
    if (flatten_scene(d, root < 0 ? d->root : root, build_bvh != 0 ? BUILD_HOST : BUILD_NONE, e->flat, err) != RT_OK) {
        g_err = err;
        delete e;
        return nullptr;
    }
    make_view(*e);
    return e;
}
void emul_scene_destroy(void* h) { delete (EmulScene*)h; }
void emul_scene_info(void* h, int32_t* n_prims, int32_t* n_nodes, int32_t* n_media, int32_t* depth) {
    EmulScene* e = (EmulScene*)h;
    *n_prims = (int)e->flat.prims.size(), *n_nodes = (int)e->flat.nodes.size(), *n_media = (int)e->flat.media.size();
    *depth = e->flat.bvh_depth;
}
void emul_scene_info4(void* h, int32_t* n_nodes4, int32_t* depth4) {
    EmulScene* e = (EmulScene*)h;
    *n_nodes4 = (int)e->flat.nodes4.size(), *depth4 = e->flat.bvh4_depth;
}
int32_t emul_prim_nodes(void* h, int32_t* out, int32_t cap) {
    EmulScene* e = (EmulScene*)h;
    for (int i = 0; i < cap && i < (int)e->flat.prim_node.size(); ++i) out[i] = e->flat.prim_node[i];
    return (int)e->flat.prim_node.size();
}

void emul_intersect_batch(void* h, int32_t mode, const float* rays, int64_t n, RtHit* out) {
    EmulScene* e = (EmulScene*)h;
    for (int64_t i = 0; i < n; ++i) intersect_query(e->view, mode, rays + 8 * i, out[i]);
}
void emul_intersect_batch_at(void* h, int32_t mode, float time, const float* rays, int64_t n, RtHit* out) {
    EmulScene* e = (EmulScene*)h;
    for (int64_t i = 0; i < n; ++i) intersect_query(e->view, mode, rays + 8 * i, out[i], time);
}

void emul_scatter_batch(void* h, const RtScatterIn* in, int64_t n, RtScatterOut* out) {
    EmulScene* e = (EmulScene*)h;
    for (int64_t i = 0; i < n; ++i) scatter_query(e->view, in[i], out[i]);
}

void emul_texture_value_batch(void* h, int32_t tex, const float* uvp, int64_t n, float* rgb) {
    EmulScene* e = (EmulScene*)h;
    for (int64_t i = 0; i < n; ++i) {
        const float* q = uvp + 5 * i;
        V3 c = texture_value(e->view, tex, q[0], q[1], v3(q[2], q[3], q[4]));
        rgb[3 * i] = c.x, rgb[3 * i + 1] = c.y, rgb[3 * i + 2] = c.z;
    }
}

static DRenderParams make_params(const RtParams* p, int spi, int chunks) {
    DRenderParams P{};
    P.width = p->width, P.height = p->height, P.max_depth = p->max_depth;
    P.sample_begin = p->sample_begin, P.samples_per_item = spi, P.items_per_pixel = chunks;
    P.tiles_x = (p->width + 7) / 8, P.tiles_y = (p->height + 3) / 4;
    P.seed_lo = (uint32_t)p->seed, P.seed_hi = (uint32_t)(p->seed >> 32);
    P.inv_wm1 = 1.0f / ((float)p->width - 1.0f), P.inv_hm1 = 1.0f / ((float)p->height - 1.0f);
    return P;
}

void emul_generate_rays(const RtCamera* cam, const RtParams* p, const int32_t* pixel, const int32_t* sample, int64_t n, float* rays, float* us) {
    DCamera dc;
    make_camera(*cam, dc);
    DRenderParams P = make_params(p, 1, 1);
    for (int64_t i = 0; i < n; ++i) camera_query(dc, P, pixel[i], sample[i], rays + 6 * i, us + 4 * i);
}

void emul_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out); }

void emul_unit_ball(const float* u, int64_t n, float* out) {
    for (int64_t i = 0; i < n; ++i) {
        V3 p = sample_unit_ball(u[3 * i], u[3 * i + 1], u[3 * i + 2]);
        out[3 * i] = p.x, out[3 * i + 1] = p.y, out[3 * i + 2] = p.z;
    }
}

// the megakernel's per-thread routine, run row-parallel on host threads; accum = 3*W*H radiance sums in the library's
// fixed point (AccumFx: 2^-32 units)
uint64_t emul_render(void* h, const RtCamera* cam, const RtParams* p, int32_t threads, uint64_t* accum) {
    EmulScene* e = (EmulScene*)h;
    DCamera dc;
    make_camera(*cam, dc);
    int count = p->sample_count > 0 ? p->sample_count : p->samples_per_pixel;
    DRenderParams P = make_params(p, count, 1);
    std::atomic<int> next(0);
    std::atomic<uint64_t> rays(0);
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    auto worker = [&]() {
        uint64_t local = 0;
        for (;;) {
            int j = next.fetch_add(1);
            if (j >= p->height) break;
            for (int i = 0; i < p->width; ++i) {
                AccumFx sum[3];
                uint32_t nr = 0;
                integrate_item(e->view, dc, P, i, j, p->sample_begin, count, sum, nr);
                local += nr;
                for (int k = 0; k < 3; ++k) accum[3 * ((size_t)j * p->width + i) + k] = sum[k];
            }
        }
        rays += local;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
    return rays.load();
}

uint32_t emul_clear_media(void* h) { return ((EmulScene*)h)->flat.clear_media; }
uint32_t emul_scene_features(void* h) { return ((EmulScene*)h)->flat.features; }

// The wavefront slot functions (wf_init_pixel_sample, closest hit, wf_shade) driven path by path on host threads, as the
// persistent kernel drives them — with `use_chain`, paths inside a clear medium advance with wf_chain_step instead of a
// surface search + wf_shade, as in the kernel's chain phase.  counters[0] = rays, counters[1] = chain steps.
void emul_render_slots(void* h, const RtCamera* cam, const RtParams* p, int32_t threads, int32_t use_chain, uint64_t* accum, uint64_t* counters) {
    EmulScene* e = (EmulScene*)h;
    const DSceneView& S = e->view;
    DCamera dc;
    make_camera(*cam, dc);
    int count = p->sample_count > 0 ? p->sample_count : p->samples_per_pixel;
    DRenderParams P = make_params(p, 1, 1);
    std::atomic<int> next(0);
    std::atomic<uint64_t> rays(0), steps(0);
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    auto worker = [&]() {
        uint64_t local_rays = 0, local_steps = 0;
        for (;;) {
            int j = next.fetch_add(1);
            if (j >= p->height) break;
            for (int i = 0; i < p->width; ++i) {
                AccumFx sum[3] = {0, 0, 0};
                const uint32_t pixel = (uint32_t)(j * p->width + i);
                for (int k = 0; k < count && p->max_depth > 0; ++k) {
                    WfSlot s;
                    wf_init_pixel_sample(S, dc, P, pixel, (uint32_t)(p->sample_begin + k), s);
                    bool chain = false;
                    for (;;) {
                        local_rays += 1;
                        if (chain) {
                            local_steps += 1;
                            const int rc = wf_chain_step(S, P, s);
                            if (rc == 0) break;  // Color::ZERO
                            chain = rc == 1;
                            continue;
                        }
                        Ray r;
                        r.o = v3(s.A.x, s.A.y, s.A.z), r.d = v3(s.B.x, s.B.y, s.B.z);
                        const uint32_t flags = as_uint(s.B.w);
                        float t = s.D.x;
                        int prim, face;
                        closest_hit(S, r, RTB_T_MIN, t, (int)as_uint(s.D.z), (int)((flags >> WF_FACE_SHIFT) & 7u), t, prim, face, time_of_flags(flags));
                        if (prim >= 0) s.D.x = t, s.D.y = as_float((uint32_t)(prim | (face << 24)));
                        const int code = (int)as_uint(s.D.y);
                        V3 radiance;
                        if (!wf_shade(S, P, s, radiance)) {
                            sum[0] += radiance_fixed(radiance.x), sum[1] += radiance_fixed(radiance.y), sum[2] += radiance_fixed(radiance.z);
                            break;
                        }
                        chain = use_chain != 0 && wf_chain_eligible(S, code, s);
                    }
                }
                for (int k = 0; k < 3; ++k) accum[3 * (size_t)pixel + k] = sum[k];
            }
        }
        rays += local_rays, steps += local_steps;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
    counters[0] = rays.load(), counters[1] = steps.load();
}

}  // extern "C"
