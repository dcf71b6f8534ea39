// Wavefront pipeline, warp-autonomous shared-memory form (RT_PIPELINE_WAVEFRONT_SMEM; an evaluated alternative).
//
// The generate -> extend -> shade wavefront of BASELINE.json's north_star, with the path pool and the queues of
// every warp kept in SHARED memory (B200: 227 KB per SM) instead of global memory:
//
//   * a warp owns an 8x4 pixel tile and a pool of WA_SLOTS in-flight paths for it (SoA in shared memory);
//   * "ready" slots hold a ray waiting to be traced, "pending" slots hold a hit waiting to be shaded, binned by
//     material class; both queues are byte stacks in shared memory, appended to with warp-ballot + prefix-sum ranks;
//   * extend: each lane parks one traversal (ray, closest hit, short stack) in registers; the warp votes between an
//     inner-node loop and a leaf step so that its lanes execute the same kind of work; finished lanes are retired
//     and refilled from the ready stack in batches;
//   * shade: as soon as a class has 32 pending hits the warp shades them as one full batch (per-material code is
//     warp-uniform); the new rays (scattered, or regenerated camera rays when a path ends) go back on the ready
//     stack.  The parked traversals of the lanes simply wait in registers meanwhile.
//
// Nothing but the BVH/primitive fetches (L1-resident) and one float RED per pixel channel per work item touches
// global memory; there are no global queues, no global atomics on the path and no kernel boundaries between stages.
// It replaces Renderer::render / render_pixel / trace_internal (src/raytrace.rs:172-198, :79-101).
#include <cuda_runtime.h>

#include <algorithm>

#include "rt_device.cuh"
#include "scene_internal.h"

namespace rtb {

#ifndef WA_SLOTS
#define WA_SLOTS 192  // paths in flight per warp (multiple of 32, at most 256: slot ids are bytes)
#endif
#define WA_K (WA_SLOTS / 32)
#define WA_WARPS 4
#define WA_THREADS (WA_WARPS * 32)
#define WA_DONE RTB_TRAVERSAL_DONE
#define WA_REFILL 8   // idle lanes before the warp leaves the traversal loop to refill them
#define WA_FLUSH 16   // idle lanes (with nothing ready) before a partial class batch is shaded
#define WA_MIN_DESCEND 8
enum { F_OX, F_OY, F_OZ, F_DX, F_DY, F_DZ, F_BR, F_BG, F_BB, F_T, F_CODE, F_ORG, F_FLAGS, F_SAMPLE, WA_FIELDS };
// pending classes: everything cheap shades together (the material switch diverges over a few dozen instructions,
// the RNG / media / state work around it is common); only textured hits (Perlin, image) wait for their own batch
enum { WA_CLS_PLAIN = 0, WA_CLS_TEXTURED = 1, WA_NCLS = 2 };

struct alignas(16) WarpPool {
    float f[WA_FIELDS][WA_SLOTS];
    float accum[3][32];
    unsigned char ready[WA_SLOTS];
    unsigned char pend[WA_NCLS][WA_SLOTS];
};

struct WaJob {
    int tiles_x, n_tiles;
    int n_items;
    int samples_per_item;  // sample indices [chunk * spi, ...) of the launch belong to chunk
    int total_samples;
};

__device__ __forceinline__ void pool_load(const WarpPool& p, unsigned int s, uint32_t pixel, WfSlot& o) {
    o.A = f4(p.f[F_OX][s], p.f[F_OY][s], p.f[F_OZ][s], __uint_as_float(pixel));
    o.B = f4(p.f[F_DX][s], p.f[F_DY][s], p.f[F_DZ][s], p.f[F_FLAGS][s]);
    o.C = f4(p.f[F_BR][s], p.f[F_BG][s], p.f[F_BB][s], p.f[F_SAMPLE][s]);
    o.D = f4(p.f[F_T][s], p.f[F_CODE][s], p.f[F_ORG][s], 0.f);
}
__device__ __forceinline__ void pool_store(WarpPool& p, unsigned int s, const WfSlot& o) {
    p.f[F_OX][s] = o.A.x, p.f[F_OY][s] = o.A.y, p.f[F_OZ][s] = o.A.z;
    p.f[F_DX][s] = o.B.x, p.f[F_DY][s] = o.B.y, p.f[F_DZ][s] = o.B.z, p.f[F_FLAGS][s] = o.B.w;
    p.f[F_BR][s] = o.C.x, p.f[F_BG][s] = o.C.y, p.f[F_BB][s] = o.C.z, p.f[F_SAMPLE][s] = o.C.w;
    p.f[F_T][s] = o.D.x, p.f[F_CODE][s] = o.D.y, p.f[F_ORG][s] = o.D.z;
}

__global__ void __launch_bounds__(WA_THREADS) warpfront_kernel(DSceneView S, DCamera cam, DRenderParams P, WaJob job, float* __restrict__ accum,
                                                               unsigned long long* __restrict__ rays_out, unsigned int* __restrict__ item_counter) {
    extern __shared__ float4 wa_smem[];
    const unsigned int lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned int lt_mask = (1u << lane) - 1u;
    WarpPool& pool = reinterpret_cast<WarpPool*>(wa_smem)[warp];
    const int root_link = S.n_prims > 0 ? (int)as_uint(ld4(S.nodes).w) : WA_DONE;
    unsigned int n_rays = 0;

    for (;;) {
        // ---------------------------------------------------------------- next work item: tile x run of samples
        unsigned int item = 0;
        if (lane == 0) item = atomicAdd(item_counter, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= (unsigned int)job.n_items) break;
        const int tile = (int)(item % (unsigned int)job.n_tiles), chunk = (int)(item / (unsigned int)job.n_tiles);
        const int tile_x = tile % job.tiles_x, tile_y = tile / job.tiles_x;
        const int first_sample = chunk * job.samples_per_item;
        const int n_samp = min(job.samples_per_item, job.total_samples - first_sample);
        const int sample0 = P.sample_begin + first_sample;  // global index of the item's first sample

        // lane-parked traversal
        StackEntry stack[RTB_BVH_STACK];
        int sp = 0, cur = WA_DONE;
        bool has = false;
        unsigned int my_slot = 0;
        Ray r;
        r.o = r.d = v3(0.f, 0.f, 0.f);
        NodeRay nr;
    nr.inv = nr.noi = v3(0.f, 0.f, 0.f), nr.pad = 0.f;
        float t_best = RTB_INF;
        int prim_best = -1, face_best = 0, mat_best = 0, origin_prim = -1, origin_face = 0, code_in = -1;
        // warp-uniform queue state
        unsigned int n_ready = 0, alive = 0;
        unsigned int n_pend[WA_NCLS] = {0u, 0u};

        // ---------------------------------------------------------------- generate: first WA_K samples of every pixel
        {
            const int px = tile_x * 8 + (int)(lane & 7u), py = tile_y * 4 + (int)(lane >> 3);
            const bool pixel_ok = px < P.width && py < P.height;
            const uint32_t pixel = (uint32_t)(py * P.width + px);
            pool.accum[0][lane] = 0.f, pool.accum[1][lane] = 0.f, pool.accum[2][lane] = 0.f;
#pragma unroll 1
            for (int k = 0; k < WA_K; ++k) {
                const bool ok = pixel_ok && k < n_samp;
                const unsigned int s = (unsigned int)k * 32u + lane;
                if (ok) {
                    WfSlot slot;
                    wf_init_pixel_sample(S, cam, P, pixel, (uint32_t)(sample0 + k), slot);
                    pool_store(pool, s, slot);
                }
                unsigned int m = __ballot_sync(0xffffffffu, ok);
                if (ok) pool.ready[n_ready + __popc(m & lt_mask)] = (unsigned char)s;
                n_ready += (unsigned int)__popc(m);
            }
            alive = n_ready;
            __syncwarp();
        }

        // ---------------------------------------------------------------- scheduler
        while (alive > 0u) {
            // (1) retire finished traversals: final hit -> slot, slot -> pending stack of its class
            {
                const bool fin = has && cur == WA_DONE;
                const unsigned int fin_mask = __ballot_sync(0xffffffffu, fin);
                if (fin_mask) {
                    int cls = -1;
                    if (fin) {
                        int code;
                        int c7 = wf_classify(S, prim_best, face_best, mat_best, code_in, code);
                        pool.f[F_T][my_slot] = t_best;
                        pool.f[F_CODE][my_slot] = __int_as_float(code);
                        cls = c7 == WF_TEXTURED ? WA_CLS_TEXTURED : WA_CLS_PLAIN;
                        has = false;
                    }
#pragma unroll
                    for (int c = 0; c < WA_NCLS; ++c) {
                        unsigned int m = __ballot_sync(0xffffffffu, cls == c);
                        if (cls == c) pool.pend[c][n_pend[c] + __popc(m & lt_mask)] = (unsigned char)my_slot;
                        n_pend[c] += (unsigned int)__popc(m);
                    }
                    __syncwarp();
                }
            }
            unsigned int busy = __ballot_sync(0xffffffffu, has);
            // (2) refill idle lanes from the ready stack
            if (n_ready > 0u && (32 - __popc(busy) >= WA_REFILL || busy == 0u)) {
                const unsigned int n_empty = 32u - (unsigned int)__popc(busy);
                const unsigned int take = min(n_empty, n_ready);
                const unsigned int rank = (unsigned int)__popc(~busy & lt_mask);
                if (!has && rank < take) {
                    my_slot = pool.ready[n_ready - take + rank];
                    r.o = v3(pool.f[F_OX][my_slot], pool.f[F_OY][my_slot], pool.f[F_OZ][my_slot]);
                    r.d = v3(pool.f[F_DX][my_slot], pool.f[F_DY][my_slot], pool.f[F_DZ][my_slot]);
                    t_best = pool.f[F_T][my_slot], code_in = __float_as_int(pool.f[F_CODE][my_slot]);
                    origin_prim = __float_as_int(pool.f[F_ORG][my_slot]);
                    origin_face = (int)((__float_as_uint(pool.f[F_FLAGS][my_slot]) >> WF_FACE_SHIFT) & 7u);
                    nr = node_ray(r);
                    prim_best = -1, face_best = 0;
                    sp = 0, cur = root_link;
                    has = true;
                    n_rays += 1u;
                }
                n_ready -= take;
                busy = __ballot_sync(0xffffffffu, has);
            }
            // (3) shade: a full batch of some class, or a partial one when the warp would otherwise starve
            {
                int c_pick = -1;
                unsigned int n_take = 0;
#pragma unroll
                for (int c = 0; c < WA_NCLS; ++c)
                    if (c_pick < 0 && n_pend[c] >= 32u) c_pick = c, n_take = 32u;
                if (c_pick < 0 && n_ready == 0u && 32 - __popc(busy) >= WA_FLUSH) {
                    unsigned int best = 0;
#pragma unroll
                    for (int c = 0; c < WA_NCLS; ++c)
                        if (n_pend[c] > best) best = n_pend[c], c_pick = c;
                    n_take = min(best, 32u);
                }
                if (c_pick >= 0) {
                    unsigned int base = 0;
#pragma unroll
                    for (int c = 0; c < WA_NCLS; ++c)
                        if (c == c_pick) base = n_pend[c] - n_take, n_pend[c] = base;
                    const bool act = lane < n_take;
                    unsigned int s = 0;
                    bool lives = false;
                    WfSlot slot;
                    if (act) {
                        s = pool.pend[c_pick][base + lane];
                        const unsigned int pl = s & 31u;
                        const uint32_t pixel = (uint32_t)((tile_y * 4 + (int)(pl >> 3)) * P.width + tile_x * 8 + (int)(pl & 7u));
                        pool_load(pool, s, pixel, slot);
                        V3 radiance;
                        lives = wf_shade(S, P, slot, radiance);
                        if (!lives) {
                            // path ended: deposit, then start this slot's next sample (stride WA_K) if the item has one
                            if (radiance.x != 0.f) atomicAdd(&pool.accum[0][pl], radiance.x);
                            if (radiance.y != 0.f) atomicAdd(&pool.accum[1][pl], radiance.y);
                            if (radiance.z != 0.f) atomicAdd(&pool.accum[2][pl], radiance.z);
                            const int next = (int)__float_as_uint(slot.C.w) - sample0 + WA_K;
                            if (next < n_samp) {
                                wf_init_pixel_sample(S, cam, P, pixel, (uint32_t)(sample0 + next), slot);
                                lives = true;
                            }
                        }
                        if (lives) pool_store(pool, s, slot);
                    }
                    const unsigned int m = __ballot_sync(0xffffffffu, lives);
                    if (lives) pool.ready[n_ready + __popc(m & lt_mask)] = (unsigned char)s;
                    n_ready += (unsigned int)__popc(m);
                    alive -= n_take - (unsigned int)__popc(m);
                    __syncwarp();
                    continue;
                }
            }
            if (busy == 0u) continue;  // nothing parked: the next pass refills or flushes

            // (4) extend: traverse until enough lanes are idle again
            for (;;) {
                // inner nodes: lanes leave the loop when they reach a leaf or run out of nodes
                while (has && cur >= 0) {
                    const char* base = reinterpret_cast<const char*>(S.nodes + cur);
                    float4 l0 = ld4(base), l1 = ld4(base + 16), r0 = ld4(base + 32), r1 = ld4(base + 48);
                    float tl, tr;
                    bool hl = slab_node(l0, l1, nr, RTB_T_MIN, t_best, tl);
                    bool hr = slab_node(r0, r1, nr, RTB_T_MIN, t_best, tr);
                    int ll = (int)as_uint(l0.w), lr = (int)as_uint(r0.w);
                    if (hl && hr) {
                        bool left_first = tl <= tr;
                        stack[sp].node = left_first ? lr : ll, stack[sp].tn = left_first ? tr : tl;
                        sp += 1;
                        cur = left_first ? ll : lr;
                    } else if (hl) {
                        cur = ll;
                    } else if (hr) {
                        cur = lr;
                    } else {
                        cur = stack_pop(stack, sp, t_best, nr.pad);
                    }
                    if (__popc(__activemask()) < WA_MIN_DESCEND) break;
                }
                __syncwarp();
                // one leaf: every primitive of it, then pop
                if (has && cur < 0 && cur != WA_DONE) {
                    int v = ~cur;
                    int first = v & 0xFFFFFF, count = v >> 24;
                    for (int i = first; i < first + count; ++i) {
                        PrimRec p = load_prim(S.prims + i);
                        float t;
                        int face;
                        if (hit_prim(S, p, r, RTB_T_MIN, t_best, i == origin_prim, origin_face, t, face))
                            t_best = t, prim_best = i, face_best = face, mat_best = p.mat;
                    }
                    cur = stack_pop(stack, sp, t_best, nr.pad);
                }
                __syncwarp();
                const unsigned int working = __ballot_sync(0xffffffffu, has && cur != WA_DONE);
                if (working == 0u) break;
                const int idle = 32 - __popc(working);
                if (n_ready > 0u ? idle >= WA_REFILL : idle >= WA_FLUSH) break;
            }
        }

        // ---------------------------------------------------------------- item done: one RED per pixel channel
        __syncwarp();
        {
            const int px = tile_x * 8 + (int)(lane & 7u), py = tile_y * 4 + (int)(lane >> 3);
            if (px < P.width && py < P.height) {
                float* dst = accum + 3 * ((size_t)py * P.width + px);
                atomicAdd(dst + 0, pool.accum[0][lane]);
                atomicAdd(dst + 1, pool.accum[1][lane]);
                atomicAdd(dst + 2, pool.accum[2][lane]);
            }
        }
        __syncwarp();
    }
    for (int off = 16; off > 0; off >>= 1) n_rays += __shfl_down_sync(0xffffffffu, n_rays, off);
    if (lane == 0 && n_rays) atomicAdd(rays_out, (unsigned long long)n_rays);
}

// ---------------------------------------------------------------------------------------------------------------
struct WarpfrontState {
    unsigned int* d_item_counter = nullptr;
    int blocks = 0;
    size_t smem = 0;
};

static WarpfrontState* warpfront_state(RtScene* s, int* rc) {
    static_assert(WA_SLOTS % 32 == 0 && WA_SLOTS <= 256, "slot ids are bytes");
    *rc = RT_OK;
    if (s->wa) return s->wa;
    WarpfrontState* w = new WarpfrontState();
    w->smem = sizeof(WarpPool) * WA_WARPS;
    if (cudaMalloc(&w->d_item_counter, sizeof(unsigned int)) != cudaSuccess ||
        cudaFuncSetAttribute(warpfront_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w->smem) != cudaSuccess) {
        *rc = set_error(RT_ERR_CUDA, "warpfront: setup failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete w;
        return nullptr;
    }
    int per_sm = 0, sms = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, warpfront_kernel, WA_THREADS, w->smem);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device);
    w->blocks = std::max(1, per_sm) * std::max(1, sms);
    s->wa = w;
    return w;
}

void free_warpfront(RtScene* s) {
    if (!s->wa) return;
    cudaFree(s->wa->d_item_counter);
    delete s->wa;
    s->wa = nullptr;
}

bool warpfront_supports(const RtScene* s, const RtParams* p) {
    return p->max_depth <= (int)WF_DEPTH_MASK && s->flat.prims.size() < (1u << 24);
}

int launch_warpfront(RtScene* s, const DCamera& cam, const RtParams* p, int begin, int count, float* d_accum, cudaStream_t stream, RtProgressFn cb,
                     void* user, int* launches) {
    if (!warpfront_supports(s, p)) return set_error(RT_ERR_UNSUPPORTED, "wavefront pipeline: max_depth above %d or more than 2^24 primitives", WF_DEPTH_MASK);
    if (p->max_depth <= 0) return RT_OK;  // every path returns Color::ZERO at once (raytrace.rs:87-89)
    int rc;
    WarpfrontState* w = warpfront_state(s, &rc);
    if (!w) return rc;
    WaJob job;
    job.tiles_x = (p->width + 7) / 8;
    job.n_tiles = job.tiles_x * ((p->height + 3) / 4);
    int spi = p->samples_per_item > 0 ? p->samples_per_item : 32 * WA_K;  // every slot integrates ~32 paths per item
    spi = std::max(1, std::min(spi, count));
    job.samples_per_item = spi;
    // bound one launch to ~2^28 camera paths so that progress can be reported
    long long per_chunk = (long long)job.n_tiles * 32 * spi;
    int chunks_total = (count + spi - 1) / spi;
    int chunks_per_launch = (int)std::max<long long>(1, (1ll << 28) / per_chunk);
    int done = 0;
    for (int c0 = 0; c0 < chunks_total; c0 += chunks_per_launch) {
        int chunks = std::min(chunks_per_launch, chunks_total - c0);
        int samples = std::min(count - done, chunks * spi);
        DRenderParams P = device_params(p, begin + done, spi, chunks);
        job.total_samples = samples;
        job.n_items = job.n_tiles * chunks;
        CU_TRY(cudaMemsetAsync(w->d_item_counter, 0, sizeof(unsigned int), stream));
        int blocks = std::min(w->blocks, (job.n_items + WA_WARPS - 1) / WA_WARPS);
        warpfront_kernel<<<blocks, WA_THREADS, w->smem, stream>>>(s->view, cam, P, job, d_accum, s->d_rays, w->d_item_counter);
        CU_TRY(cudaGetLastError());
        *launches += 1;
        done += samples;
        if (cb) {
            CU_TRY(cudaStreamSynchronize(stream));
            cb(done, count, user);
        }
    }
    return RT_OK;
}

}  // namespace rtb
