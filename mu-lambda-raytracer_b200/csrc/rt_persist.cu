// Persistent pipeline: the wavefront stages (generate / extend / shade) run inside ONE resident kernel, with the path
// pool in registers + shared memory and warp ballots deciding which stage a warp executes next.
//
// It replaces Renderer::render / render_pixel / trace_internal (src/raytrace.rs:172-198, :79-101).  Every lane owns TWO
// camera paths: one in registers (being traversed) and one parked in shared memory (waiting to be shaded, or
// holding a fresh ray).  The warp alternates between
//   extend  : while-while BVH traversal of the register paths (32-byte nodes, 128-bit __ldg, short stack); lanes vote
//             between the inner-node loop and a leaf step, exactly like wf_extend_kernel;
//   shade   : when enough lanes hold a finished traversal (or an empty slot), those paths are shaded together —
//             scatter / emit / background, terminated paths deposit beta * radiance with float REDs and are
//             regenerated in place from a global camera-path counter (reserved in chunks, one atomic per 256 paths),
//             the media event of the new ray is pre-sampled — and become "ready" again.
// A lane whose register path finishes simply swaps it with its parked ready path and keeps traversing, so the extend
// stage runs with (nearly) full warps and the shade stage with (nearly) full warps, without any queue in global memory:
// HBM only sees the accumulation REDs.  Noise textures are evaluated warp-cooperatively (the 56 gradient terms of
// the 7-octave turbulence are spread over the lanes) instead of by one lane while 31 wait.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "rt_device.cuh"
#include "scene_internal.h"

namespace rtb {

#define PS_THREADS 128
#define PS_DONE ((int)0x80000000)
#define PS_MIN_DESCEND 6   // lanes that must still be descending inner nodes for the inner loop to keep going
#define PS_LEAVE 8         // lanes that must have finished before the warp leaves the traversal loop to swap paths
#define PS_WORK 24         // lanes with shading / regeneration work pending that trigger a shade phase ...
#define PS_STALL 14        // ... or lanes that cannot traverse at all (both of their paths wait for the shade phase)
#define PS_MIN_BLOCKS 5    // resident blocks per SM the kernel is compiled for: 96 registers, no spills (6/7/8 blocks spill and measured 4/6/11 % slower)
#define PS_CHUNK 256u      // camera paths a warp reserves per atomic

enum { ST_EMPTY = 0, ST_TRAV = 1, ST_DONE = 2, ST_READY = 3 };

// scheduling thresholds (defaults = the PS_* constants; RT_PS_WORK / RT_PS_STALL / RT_PS_LEAVE / RT_PS_DESCEND override them for sweeps)
struct PsTune {
    int work, stall, leave, descend;
};

struct PsCounters {
    unsigned long long next_path, total_paths, rays;
};

struct PersistState {
    PsCounters* ctr = nullptr;
    int blocks = 0;
};

// one camera path of a lane.  hit: -1 miss, prim | face << 24 surface, WF_MEDIUM | m medium (as in WfSlot.D.y);
// flags: depth left | origin face << 16 (as in WfSlot.B.w)
struct LanePath {
    float ox, oy, oz, dx, dy, dz;
    float br, bg, bb;
    float t_best;
    uint32_t pixel, sample, flags;
    int origin_prim, hit, cur, sp, status;
};
#define PS_WORDS 18

__device__ __forceinline__ void park_swap(float (*park)[PS_THREADS], LanePath& p) {
    const unsigned int t = threadIdx.x;
#define SWAPF(k, field)           \
    {                             \
        float tmp = park[k][t];   \
        park[k][t] = p.field;     \
        p.field = tmp;            \
    }
#define SWAPI(k, field)                                  \
    {                                                    \
        float tmp = park[k][t];                          \
        park[k][t] = __int_as_float((int)p.field);       \
        p.field = (decltype(p.field))__float_as_int(tmp); \
    }
    SWAPF(0, ox) SWAPF(1, oy) SWAPF(2, oz) SWAPF(3, dx) SWAPF(4, dy) SWAPF(5, dz) SWAPF(6, br) SWAPF(7, bg) SWAPF(8, bb) SWAPF(9, t_best)
    SWAPI(10, pixel) SWAPI(11, sample) SWAPI(12, flags) SWAPI(13, origin_prim) SWAPI(14, hit) SWAPI(15, cur) SWAPI(16, sp) SWAPI(17, status)
#undef SWAPF
#undef SWAPI
}

__device__ __forceinline__ void to_slot(const LanePath& p, WfSlot& s) {
    s.A = f4(p.ox, p.oy, p.oz, __uint_as_float(p.pixel));
    s.B = f4(p.dx, p.dy, p.dz, __uint_as_float(p.flags));
    s.C = f4(p.br, p.bg, p.bb, __uint_as_float(p.sample));
    s.D = f4(p.t_best, __int_as_float(p.hit), __int_as_float(p.origin_prim), 0.f);
}
__device__ __forceinline__ void from_slot(const WfSlot& s, LanePath& p) {
    p.ox = s.A.x, p.oy = s.A.y, p.oz = s.A.z, p.pixel = __float_as_uint(s.A.w);
    p.dx = s.B.x, p.dy = s.B.y, p.dz = s.B.z, p.flags = __float_as_uint(s.B.w);
    p.br = s.C.x, p.bg = s.C.y, p.bb = s.C.z, p.sample = __float_as_uint(s.C.w);
    p.t_best = s.D.x, p.hit = __float_as_int(s.D.y), p.origin_prim = __float_as_int(s.D.z);
}

// Perlin::turbulence (textures.rs:76-88, depth 7) of one point, computed by the whole warp: term = (octave, corner),
// 56 terms over 32 lanes, then a butterfly sum.  `p` must be warp-uniform.  Every lane returns the result.
__device__ __forceinline__ float warp_turbulence(const float* __restrict__ vec, const unsigned short* __restrict__ perm, V3 p) {
    const int lane = (int)(threadIdx.x & 31u);
    float acc = 0.0f;
#pragma unroll
    for (int base = 0; base < 56; base += 32) {
        const int term = base + lane;
        if (term < 56) {
            const int oct = term >> 3, di = (term >> 2) & 1, dj = (term >> 1) & 1, dk = term & 1;
            const float scale = (float)(1 << oct);
            const float qx = p.x * scale, qy = p.y * scale, qz = p.z * scale;  // exact: the reference doubles p per octave
            const float fx = floorf(qx), fy = floorf(qy), fz = floorf(qz);
            const float u = qx - fx, v = qy - fy, w = qz - fz;
            const int i = (int)fx, j = (int)fy, k = (int)fz;
            const float uu = u * u * (3.0f - 2.0f * u), vv = v * v * (3.0f - 2.0f * v), ww = w * w * (3.0f - 2.0f * w);
            const int idx = perm[(i + di) & 1023] ^ perm[1024 + ((j + dj) & 1023)] ^ perm[2048 + ((k + dk) & 1023)];
            const float4 g = ld4(vec + 4 * idx);
            const float wx = u - (float)di, wy = v - (float)dj, wz = w - (float)dk;
            const float bi = di ? uu : 1.0f - uu, bj = dj ? vv : 1.0f - vv, bk = dk ? ww : 1.0f - ww;
            acc += (1.0f / scale) * (bi * bj * bk * (wx * g.x + wy * g.y + wz * g.z));
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    return fabsf(acc);
}

__global__ void ps_reset_kernel(PsCounters* c, unsigned long long total) { c->next_path = 0, c->total_paths = total, c->rays = 0; }

// phase statistics of a debug launch (RT_PS_STATS=1): warp-level step counts and the lanes that were useful in them
enum { PSS_SHADE_PHASES, PSS_SHADE_ACT, PSS_SHADE_DONE, PSS_SHADE_ONPARK, PSS_EXT_PHASES, PSS_INNER_ITERS, PSS_INNER_LANES, PSS_LEAF_STEPS,
       PSS_LEAF_LANES, PSS_LEAF_PRIMS, PSS_EXT_ROUNDS, PSS_EXT_TRAV_LANES, PSS_NOISE, PSS_COUNT };

template <bool STATS>
__global__ void __launch_bounds__(PS_THREADS, STATS ? 4 : PS_MIN_BLOCKS) persist_kernel(DSceneView S, DCamera cam, DRenderParams P, PsCounters* __restrict__ ctr,
                                                             float* __restrict__ accum, unsigned long long* __restrict__ rays_out,
                                                             unsigned int chunk_size, unsigned long long* __restrict__ stats, PsTune tune) {
    unsigned int st_[PSS_COUNT];
    if (STATS)
        for (int k = 0; k < PSS_COUNT; ++k) st_[k] = 0u;
#define PS_STAT(k, v)                              \
    if (STATS) {                                   \
        const unsigned int v_ = (unsigned int)(v); \
        if (lane == 0) st_[k] += v_;               \
    }
    __shared__ float park[PS_WORDS][PS_THREADS];
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int lt_mask = (1u << lane) - 1u;
    const int root_link = S.n_prims > 0 ? (int)as_uint(ld4(S.nodes).w) : PS_DONE;
    const unsigned long long total = ctr->total_paths;
    const unsigned int npix = (unsigned int)P.width * (unsigned int)P.height;

    LanePath p;  // the register path
    p.ox = p.oy = p.oz = p.dx = p.dy = p.dz = 0.f, p.br = p.bg = p.bb = 0.f, p.t_best = 0.f;
    p.pixel = p.sample = p.flags = 0u, p.origin_prim = -1, p.hit = -1, p.cur = PS_DONE, p.sp = 0, p.status = ST_EMPTY;
    park[17][threadIdx.x] = __int_as_float(ST_EMPTY);
    int park_status = ST_EMPTY;  // mirror of park[17] in a register
    int stack[RTB_BVH_STACK];
    NodeRay nr;
    nr.inv = nr.noi = v3(0.f, 0.f, 0.f), nr.pad = 0.f;
    // warp-uniform reservation of camera-path numbers
    unsigned long long chunk_next = 0, chunk_end = 0;
    uint32_t chunk_pixel = 0, chunk_sample = 0;  // (pixel, sample) of path chunk_next
    bool exhausted = false;
    unsigned int n_rays = 0;

    for (;;) {
        // ---- (1) a lane whose register path cannot traverse takes its parked ready path; a ready path starts traversing
        if (p.status != ST_TRAV) {
            if (p.status != ST_READY && park_status == ST_READY) {
                const int st = p.status;
                park_swap(park, p);  // p.status <- READY, park <- st
                park_status = st;
            }
            if (p.status == ST_READY) {
                p.status = ST_TRAV;
                p.cur = root_link, p.sp = 0;
                Ray r0;
                r0.o = v3(p.ox, p.oy, p.oz), r0.d = v3(p.dx, p.dy, p.dz);
                nr = node_ray(r0);
            }
        }
        const bool trav = p.status == ST_TRAV;
        const bool p_work = p.status == ST_DONE || (p.status == ST_EMPTY && !exhausted);
        const bool k_work = park_status == ST_DONE || (park_status == ST_EMPTY && !exhausted);
        const unsigned int m_trav = __ballot_sync(0xffffffffu, trav);
        const unsigned int m_work = __ballot_sync(0xffffffffu, p_work || k_work);
        if ((m_trav | m_work) == 0u) break;

        if (m_work && (m_trav == 0u || __popc(m_work) >= tune.work || __popc(m_work & ~m_trav) >= tune.stall)) {
            // ================================================================ shade / regenerate phase
            const bool act = p_work || k_work;
            const bool on_park = act && !p_work;  // the register path is busy traversing: work on the parked one
            PS_STAT(PSS_SHADE_PHASES, 1)
            PS_STAT(PSS_SHADE_ACT, __popc(m_work))
            PS_STAT(PSS_SHADE_ONPARK, __popc(__ballot_sync(0xffffffffu, on_park)))
            if (on_park) {  // the traversing path waits in the parked record (its stack and node-test constants stay where they are)
                const int st = p.status;
                park_swap(park, p);
                park_status = st;
            }
            bool alive = false;
            V3 radiance = v3(0.f, 0.f, 0.f);
            NoiseReq req;
            req.tex = -1, req.p = v3(0.f, 0.f, 0.f);
            int segment_next = 0;
            WfSlot s;
            const bool shade = act && p.status == ST_DONE;
            PS_STAT(PSS_SHADE_DONE, __popc(__ballot_sync(0xffffffffu, shade)))
            if (shade) {
                n_rays += 1u;
                to_slot(p, s);
                alive = wf_shade_core(S, P, s, radiance, &req, segment_next);
            }
            // noise textures: one point at a time, all lanes on its 56 gradient terms
            {
                unsigned int m_noise = __ballot_sync(0xffffffffu, req.tex >= 0);
                while (m_noise) {
                    PS_STAT(PSS_NOISE, 1)
                    const int src = __ffs((int)m_noise) - 1;
                    m_noise &= m_noise - 1u;
                    const int tex = __shfl_sync(0xffffffffu, req.tex, src);
                    V3 q;
                    q.x = __shfl_sync(0xffffffffu, req.p.x, src), q.y = __shfl_sync(0xffffffffu, req.p.y, src), q.z = __shfl_sync(0xffffffffu, req.p.z, src);
                    const DTexture& T = S.texs[tex];
                    const float* vec = S.perlin_vec + (size_t)T.a * RTB_PERLIN_POINTS * 4;
                    const unsigned short* perm = S.perlin_perm + (size_t)T.a * RTB_PERLIN_POINTS * 3;
                    const float turb = warp_turbulence(vec, perm, T.scale * q);
                    if ((int)lane == src) {
                        const float g = noise_value(T, q, turb);
                        if (alive) s.C.x *= g, s.C.y *= g, s.C.z *= g;
                        else radiance = g * radiance;
                    }
                }
            }
            if (shade && !alive) {  // the path ended: beta * (emission | background | 0) -> its pixel
                float* dst = accum + 3 * (size_t)p.pixel;
                if (radiance.x != 0.f) atomicAdd(dst + 0, radiance.x);
                if (radiance.y != 0.f) atomicAdd(dst + 1, radiance.y);
                if (radiance.z != 0.f) atomicAdd(dst + 2, radiance.z);
            }
            // regenerate: the next camera paths of the job, handed out in reservation order
            bool need = act && !alive;
            unsigned int m_need = __ballot_sync(0xffffffffu, need);
            while (m_need) {
                const unsigned int avail = (unsigned int)(chunk_end - chunk_next);
                if (avail == 0u) {
                    if (exhausted) break;
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(&ctr->next_path, (unsigned long long)chunk_size);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    chunk_next = base < total ? base : total;
                    chunk_end = base + chunk_size < total ? base + chunk_size : total;
                    if (chunk_next == chunk_end) {
                        exhausted = true;
                        break;
                    }
                    chunk_pixel = (uint32_t)(chunk_next % npix), chunk_sample = (uint32_t)P.sample_begin + (uint32_t)(chunk_next / npix);
                    continue;
                }
                const unsigned int want = (unsigned int)__popc(m_need);
                const unsigned int take = want < avail ? want : avail;
                const unsigned int rank = (unsigned int)__popc(m_need & lt_mask);
                if (need && rank < take) {
                    uint32_t pixel = chunk_pixel + rank, sample = chunk_sample;
                    while (pixel >= npix) pixel -= npix, sample += 1u;
                    wf_init_camera(cam, P, pixel, sample, s);
                    segment_next = 0;
                    alive = true, need = false;
                }
                chunk_next += take;
                chunk_pixel += take;
                while (chunk_pixel >= npix) chunk_pixel -= npix, chunk_sample += 1u;
                m_need = __ballot_sync(0xffffffffu, need);
            }
            // survivors and fresh camera paths alike: media event of the new ray, then "ready"
            if (act) {
                if (alive) {
                    wf_presample_media(S, P, s, segment_next);
                    from_slot(s, p);
                    p.status = ST_READY;
                } else {
                    p.status = ST_EMPTY;
                }
            }
            if (on_park) {  // the traversing path comes back
                const int st = p.status;
                park_swap(park, p);
                park_status = st;
            }
            continue;
        }

        // ==================================================================== extend phase
        Ray r;
        r.o = v3(p.ox, p.oy, p.oz), r.d = v3(p.dx, p.dy, p.dz);
        bool has = trav;
        int cur = p.cur, sp = p.sp;
        float t_best = p.t_best;
        PS_STAT(PSS_EXT_PHASES, 1)
        for (;;) {
            PS_STAT(PSS_EXT_ROUNDS, 1)
            PS_STAT(PSS_EXT_TRAV_LANES, __popc(__ballot_sync(0xffffffffu, has)))
            // (a) inner nodes: lanes leave the loop when they reach a leaf or run out of nodes
            while (has && cur >= 0) {
                if (STATS) {
                    const unsigned int am = __activemask();
                    if ((int)lane == __ffs((int)am) - 1) st_[PSS_INNER_ITERS] += 1u, st_[PSS_INNER_LANES] += (unsigned int)__popc(am);
                }
                const char* base = reinterpret_cast<const char*>(S.nodes + cur);
                float4 l0 = ld4(base), l1 = ld4(base + 16), r0 = ld4(base + 32), r1 = ld4(base + 48);
                float tl, tr;
                bool hl = slab_node(l0, l1, nr, RTB_T_MIN, t_best, tl);
                bool hr = slab_node(r0, r1, nr, RTB_T_MIN, t_best, tr);
                int ll = (int)as_uint(l0.w), lr = (int)as_uint(r0.w);
                if (hl && hr) {
                    bool left_first = tl <= tr;
                    stack[sp++] = left_first ? lr : ll;
                    cur = left_first ? ll : lr;
                } else if (hl) {
                    cur = ll;
                } else if (hr) {
                    cur = lr;
                } else {
                    cur = sp > 0 ? stack[--sp] : PS_DONE;
                }
                if (__popc(__activemask()) < tune.descend) break;
            }
            __syncwarp();
            // (b) one leaf: every primitive of it, then pop
            if (STATS) {
                const unsigned int lm = __ballot_sync(0xffffffffu, has && cur < 0 && cur != PS_DONE);
                PS_STAT(PSS_LEAF_STEPS, lm != 0u)
                PS_STAT(PSS_LEAF_LANES, __popc(lm))
            }
            if (has && cur < 0 && cur != PS_DONE) {
                int v = ~cur;
                int first = v & 0xFFFFFF, count = v >> 24;
                if (STATS) st_[PSS_LEAF_PRIMS] += (unsigned int)count;
                for (int i = first; i < first + count; ++i) {
                    PrimRec q = load_prim(S.prims + i);
                    float t;
                    int face;
                    if (hit_prim(S, q, r, RTB_T_MIN, t_best, i == p.origin_prim, (int)((p.flags >> WF_FACE_SHIFT) & 7u), t, face))
                        t_best = t, p.hit = i | (face << 24);  // closer than the pre-sampled medium event, which it replaces
                }
                cur = sp > 0 ? stack[--sp] : PS_DONE;
            }
            __syncwarp();
            if (has && cur == PS_DONE) {  // traversal finished: the hit record is final
                has = false;
                p.status = ST_DONE;
            }
            const unsigned int busy = __ballot_sync(0xffffffffu, has);
            if (busy == 0u) break;
            if (__popc(m_trav & ~busy) >= tune.leave) break;
        }
        p.cur = cur, p.sp = sp, p.t_best = t_best;
    }

    for (int off = 16; off > 0; off >>= 1) n_rays += __shfl_down_sync(0xffffffffu, n_rays, off);
    if (lane == 0 && n_rays) atomicAdd(rays_out, (unsigned long long)n_rays);
    if (STATS) {  // per-lane counters (inner iterations are counted by the first active lane, leaf primitives by every lane)
        for (int k = 0; k < PSS_COUNT; ++k) {
            unsigned int v = st_[k];
            for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
            if (lane == 0 && v) atomicAdd(stats + k, (unsigned long long)v);
        }
    }
#undef PS_STAT
}

// ---------------------------------------------------------------------------------------------------------------
void free_persist(RtScene* s) {
    if (!s->ps) return;
    if (s->ps->ctr) cudaFree(s->ps->ctr);
    delete s->ps;
    s->ps = nullptr;
}

bool persist_supports(const RtScene* s, const RtParams* p) { return p->max_depth <= WF_DEPTH_MASK && s->flat.prims.size() < (1u << 24); }

int launch_persist(RtScene* s, const DCamera& cam, const RtParams* p, int begin, int count, float* d_accum, cudaStream_t stream, RtProgressFn cb,
                   void* user, int* launches) {
    if (!persist_supports(s, p)) return set_error(RT_ERR_UNSUPPORTED, "persistent pipeline: max_depth above %d or more than 2^24 primitives", WF_DEPTH_MASK);
    if (p->max_depth <= 0) return RT_OK;  // every path returns Color::ZERO at once (raytrace.rs:87-89)
    if (!s->ps) {
        PersistState* w = new PersistState();
        s->ps = w;
        CU_TRY(cudaMalloc(&w->ctr, sizeof(PsCounters)));
        int per_sm = 0, sms = 0;
        CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, persist_kernel<false>, PS_THREADS, 0));
        CU_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device));
        w->blocks = std::max(1, per_sm) * std::max(1, sms);  // persistent: exactly what is co-resident
    }
    PersistState* w = s->ps;
    const unsigned long long npix = (unsigned long long)p->width * p->height;
    // bound one launch to ~2^29 camera paths so that progress can be reported
    const int samples_per_launch = (int)std::max<unsigned long long>(1, (1ull << 29) / npix);
    int done = 0;
    while (done < count) {
        int samples = std::min(count - done, samples_per_launch);
        unsigned long long total = npix * (unsigned long long)samples;
        DRenderParams P = device_params(p, begin + done, 1, 1);
        int blocks = (int)std::min<unsigned long long>((unsigned long long)w->blocks, (total + 2 * PS_THREADS - 1) / (2 * PS_THREADS));
        // small jobs: smaller reservations so that every warp gets work
        unsigned long long per_warp = total / ((unsigned long long)blocks * (PS_THREADS / 32) * 4ull);
        unsigned int chunk = (unsigned int)std::min<unsigned long long>(PS_CHUNK, std::max<unsigned long long>(32ull, per_warp));
        chunk = (unsigned int)std::min<unsigned long long>(chunk, std::max<unsigned long long>(1ull, npix));
        ps_reset_kernel<<<1, 1, 0, stream>>>(w->ctr, total);
        PsTune tune = {PS_WORK, PS_STALL, PS_LEAVE, PS_MIN_DESCEND};
        if (const char* e = getenv("RT_PS_WORK")) tune.work = atoi(e);
        if (const char* e = getenv("RT_PS_STALL")) tune.stall = atoi(e);
        if (const char* e = getenv("RT_PS_LEAVE")) tune.leave = atoi(e);
        if (const char* e = getenv("RT_PS_DESCEND")) tune.descend = atoi(e);
        if (getenv("RT_PS_STATS")) {  // debug: phase statistics on stderr (slower kernel; never used by the bench)
            unsigned long long* d_stats = nullptr;
            CU_TRY(cudaMalloc(&d_stats, PSS_COUNT * sizeof(unsigned long long)));
            CU_TRY(cudaMemsetAsync(d_stats, 0, PSS_COUNT * sizeof(unsigned long long), stream));
            persist_kernel<true><<<blocks, PS_THREADS, 0, stream>>>(s->view, cam, P, w->ctr, d_accum, s->d_rays, chunk, d_stats, tune);
            unsigned long long h[PSS_COUNT];
            CU_TRY(cudaMemcpyAsync(h, d_stats, sizeof h, cudaMemcpyDeviceToHost, stream));
            CU_TRY(cudaStreamSynchronize(stream));
            cudaFree(d_stats);
            static const char* names[PSS_COUNT] = {"shade_phases", "shade_act_lanes", "shade_done_lanes", "shade_onpark_lanes", "ext_phases", "inner_iters",
                                                   "inner_lanes", "leaf_steps", "leaf_lanes", "leaf_prims", "ext_rounds", "ext_trav_lanes", "noise_evals"};
            fprintf(stderr, "persist stats (%llu paths):", total);
            for (int k = 0; k < PSS_COUNT; ++k) fprintf(stderr, " %s=%llu", names[k], h[k]);
            fprintf(stderr, "\n");
        } else {
            persist_kernel<false><<<blocks, PS_THREADS, 0, stream>>>(s->view, cam, P, w->ctr, d_accum, s->d_rays, chunk, nullptr, tune);
        }
        CU_TRY(cudaGetLastError());
        *launches += 2;
        done += samples;
        if (cb) {
            CU_TRY(cudaStreamSynchronize(stream));
            cb(done, count, user);
        }
    }
    return RT_OK;
}

}  // namespace rtb
