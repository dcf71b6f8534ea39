// Wavefront pipeline: generate -> extend -> shade over a pool of in-flight paths, with queue compaction between
// the stages (BASELINE.json north_star).  It replaces the recursion of RecursiveRayTracer::trace_internal
// (src/raytrace.rs:79-101) and the per-pixel sample loop of render_pixel (:188-198) by rounds over a path pool:
//
//   extend  persistent warps pull rays from the active queue (one warp-aggregated atomic per refill), walk the
//           32-byte-node BVH with a short stack, and vote between "inner node step" and "leaf primitive step" so
//           that the lanes of a warp execute the same kind of work; finished lanes are batched, get their media
//           free-flight sample, and are appended to the queue of their material class (ballot + prefix-sum ranks).
//   shade   one thread per queued path, blocks are class-uniform (per-material shade code), survivors are compacted
//           into the next active queue; terminated paths add beta * radiance to their pixel and are regenerated in
//           place from the global path counter, so the pool stays full until the job runs out of camera paths.
//
// All counters live on the device; the host only enqueues rounds (as CUDA graphs) and looks at a counter snapshot
// one batch behind to know when to stop.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <vector>

#include "rt_device.cuh"
#include "scene_internal.h"

namespace rtb {

struct WfCounters {
    unsigned int q_count[2];               // entries in the active queue of each parity
    unsigned int q_head;                   // fetch cursor of the extend stage
    unsigned int class_count[WF_CLASSES + 1];
    unsigned int done_extend, done_shade;  // "last block out" tickets
    unsigned int rounds;
    unsigned long long next_path, total_paths, rays;
};

// per-call job description, written by wf_reset_kernel: keeps the captured graph independent of the call's parameters
struct WfJob {
    DCamera cam;
    DRenderParams P;
    AccumFx* accum;
};

struct WfView {
    WfSlot* slots;          // path pool: n_slots records of 64 B (AoS: a path is two full 32-byte sectors)
    unsigned int* q[2];     // active queues (slot indices)
    unsigned int* cq;       // WF_CLASSES class queues of n_slots entries each
    WfCounters* ctr;
    WfJob* job;
    unsigned int n_slots;
};

struct WavefrontState {
    WfView view{};
    std::vector<void*> owned;
    WfCounters* h_snapshot = nullptr;  // pinned
    cudaEvent_t snap_event = nullptr;
    int ext_blocks = 0;
    cudaGraphExec_t graph = nullptr;  // a batch of rounds; depends only on the scene and the pool
};

#define WF_EXT_THREADS 128
#define WF_EXT_WARPS (WF_EXT_THREADS / 32)
#define WF_SHADE_THREADS 256
#define WF_DONE RTB_TRAVERSAL_DONE
#define WF_REFILL 8       // lanes that must be idle before a warp stops traversing to retire / fetch rays
#define WF_MIN_DESCEND 8  // lanes that must still be descending inner nodes for the inner loop to keep going
#define WF_CHUNK 64       // queue entries a warp reserves per atomic
#define WF_BIN 64         // per-warp, per-class staging entries (flushed 32 at a time)

// streaming accesses for the path pool: it is touched once per round, keep L1 for the BVH
__device__ __forceinline__ float4 ld_stream(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float4* p, float4 v) { __stcs(p, v); }

// ---------------------------------------------------------------------------------------------------------------
__global__ void wf_reset_kernel(WfView W, unsigned int n_init, unsigned long long total, DCamera cam, DRenderParams P, AccumFx* accum) {
    W.job->cam = cam, W.job->P = P, W.job->accum = accum;
    WfCounters* c = W.ctr;
    c->q_count[0] = n_init, c->q_count[1] = 0, c->q_head = 0;
    for (int k = 0; k <= WF_CLASSES; ++k) c->class_count[k] = 0;
    c->done_extend = c->done_shade = 0, c->rounds = 0;
    c->next_path = n_init, c->total_paths = total, c->rays = 0;
}

__global__ void wf_init_kernel(DSceneView S, WfView W, DCamera cam, DRenderParams P, unsigned int n_init) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_init) return;
    WfSlot s;
    wf_init_path(S, cam, P, i, s);
    W.slots[i] = s;
    W.q[0][i] = i;
}

// ---------------------------------------------------------------------------------------------------------------
// Extend: persistent warps.  Each lane owns one ray at a time; the warp alternates between an inner-node loop
// (lanes drop out as they reach a leaf; the loop ends when fewer than WF_MIN_DESCEND lanes are still descending)
// and a leaf step, and retires / refills lanes in batches of at least WF_REFILL.
__global__ void __launch_bounds__(WF_EXT_THREADS) wf_extend_kernel(DSceneView S, WfView W, int parity) {
    __shared__ unsigned int bins[WF_EXT_WARPS][WF_CLASSES][WF_BIN];
    __shared__ unsigned int bin_count[WF_EXT_WARPS][WF_CLASSES + 1];
    WfCounters* ctr = W.ctr;
    const unsigned int qn = ctr->q_count[parity];
    const unsigned int* __restrict__ Q = W.q[parity];
    const unsigned int lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned int lt_mask = (1u << lane) - 1u;
    if (lane <= WF_CLASSES) bin_count[warp][lane] = 0;
    __syncwarp();

    StackEntry stack[RTB_BVH_STACK];
    int sp = 0;
    int cur = WF_DONE;
    bool has = false;
    bool exhausted = qn == 0;
    unsigned int chunk_next = 0, chunk_end = 0;  // this warp's reserved range of queue positions
    unsigned int slot = 0, flags = 0;
    Ray r;
    r.o = r.d = v3(0.f, 0.f, 0.f);
    NodeRay nr;
    nr.inv = nr.noi = v3(0.f, 0.f, 0.f), nr.pad = 0.f;
    float t_best = RTB_INF;
    int prim_best = -1, face_best = 0, mat_best = 0, origin_prim = -1, code_in = -1;
    const int root_link = S.n_prims > 0 ? (int)as_uint(ld4(S.nodes).w) : WF_DONE;

    for (;;) {
        bool fin = has && cur == WF_DONE;
        unsigned int fin_mask = __ballot_sync(0xffffffffu, fin);
        unsigned int trav_mask = __ballot_sync(0xffffffffu, has && cur != WF_DONE);
        unsigned int empty_mask = __ballot_sync(0xffffffffu, !has);
        // ---- retire finished rays in batches: hit record + class bin (staged in shared memory, flushed by 32)
        if (fin_mask && (__popc(fin_mask | empty_mask) >= WF_REFILL || trav_mask == 0u)) {
            int cls = -1;
            if (fin) {
                int code;
                cls = wf_classify(S, prim_best, face_best, mat_best, code_in, code);
                float2 rec;
                rec.x = t_best, rec.y = __int_as_float(code);
                *reinterpret_cast<float2*>(&W.slots[slot].D) = rec;
            }
#pragma unroll
            for (int c = 0; c < WF_CLASSES; ++c) {
                unsigned int m = __ballot_sync(0xffffffffu, cls == c);
                if (m) {
                    unsigned int n0 = bin_count[warp][c];
                    if (cls == c) bins[warp][c][n0 + __popc(m & lt_mask)] = slot;
                    unsigned int n1 = n0 + (unsigned int)__popc(m);
                    __syncwarp();
                    if (n1 >= 32u) {  // flush one full, coalesced line of slot indices
                        unsigned int base = 0;
                        if (lane == 0) base = atomicAdd(&ctr->class_count[c], 32u);
                        base = __shfl_sync(0xffffffffu, base, 0);
                        W.cq[(size_t)c * W.n_slots + base + lane] = bins[warp][c][lane];
                        unsigned int keep = lane + 32u < n1 ? bins[warp][c][lane + 32u] : 0u;
                        __syncwarp();
                        bins[warp][c][lane] = keep;
                        n1 -= 32u;
                    }
                    __syncwarp();
                    if (lane == 0) bin_count[warp][c] = n1;
                    __syncwarp();
                }
            }
            if (fin) has = false;
            empty_mask |= fin_mask;
        }
        // ---- refill empty lanes from this warp's chunk of the active queue (one atomic per WF_CHUNK rays)
        if (!exhausted && empty_mask && (__popc(empty_mask) >= WF_REFILL || trav_mask == 0u)) {
            if (chunk_next == chunk_end) {
                unsigned int base = 0;
                if (lane == 0) base = atomicAdd(&ctr->q_head, (unsigned int)WF_CHUNK);
                base = __shfl_sync(0xffffffffu, base, 0);
                chunk_next = min(base, qn), chunk_end = min(base + (unsigned int)WF_CHUNK, qn);
                if (chunk_next == chunk_end) exhausted = true;
            }
            unsigned int my = chunk_next + (unsigned int)__popc(empty_mask & lt_mask);
            if (!has && my < chunk_end) {
                slot = Q[my];
                const float4* rec = reinterpret_cast<const float4*>(W.slots + slot);
                float4 a = ld_stream(rec), b = ld_stream(rec + 1), d = ld_stream(rec + 3);
                r.o = v3(a.x, a.y, a.z), r.d = v3(b.x, b.y, b.z);
                flags = __float_as_uint(b.w);
                t_best = d.x, code_in = __float_as_int(d.y), origin_prim = __float_as_int(d.z);
                nr = node_ray(r);
                prim_best = -1, face_best = 0;
                sp = 0, cur = root_link;
                has = true;
            }
            chunk_next = min(chunk_next + (unsigned int)__popc(empty_mask), chunk_end);
        }
        if (__ballot_sync(0xffffffffu, has) == 0u) {
            if (exhausted) break;
            continue;  // chunk ran dry mid-refill: fetch the next one
        }

        // ---- traversal
        for (;;) {
            // (a) inner nodes: lanes leave the loop when they reach a leaf or run out of nodes
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
                if (__popc(__activemask()) < WF_MIN_DESCEND) break;  // too few lanes descending: let the others catch up
            }
            __syncwarp();
            // (b) one leaf: every primitive of it, then pop
            if (has && cur < 0 && cur != WF_DONE) {
                int v = ~cur;
                int first = v & 0xFFFFFF, count = v >> 24;
                for (int i = first; i < first + count; ++i) {
                    PrimRec p = load_prim(S.prims + i);
                    float t;
                    int face;
                    apply_motion(S, p, time_of_flags(flags));  // EXTENSION: moving spheres
                    if (hit_prim(S, p, r, RTB_T_MIN, t_best, i == origin_prim, (int)((flags >> WF_FACE_SHIFT) & 7u), t, face))
                        t_best = t, prim_best = i, face_best = face, mat_best = p.mat;
                }
                cur = stack_pop(stack, sp, t_best, nr.pad);
            }
            __syncwarp();
            // (c) leave for retire/refill when enough lanes are idle, or when nothing is left to traverse
            unsigned int busy = __ballot_sync(0xffffffffu, has && cur != WF_DONE);
            if (busy == 0u) break;
            if (!exhausted && 32 - __popc(busy) >= WF_REFILL) break;
        }
    }

    // ---- flush the partially filled class bins
#pragma unroll
    for (int c = 0; c < WF_CLASSES; ++c) {
        unsigned int n = bin_count[warp][c];
        if (n) {
            unsigned int base = 0;
            if (lane == 0) base = atomicAdd(&ctr->class_count[c], n);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (lane < n) W.cq[(size_t)c * W.n_slots + base + lane] = bins[warp][c][lane];
        }
    }
    // ---- last block out: account the rays of this round and rewind the fetch cursor
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned int ticket = atomicAdd(&ctr->done_extend, 1u);
        if (ticket == gridDim.x - 1) {
            ctr->done_extend = 0;
            ctr->q_head = 0;
            ctr->rays += qn;
            ctr->rounds += 1;
            __threadfence();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// block-wide exclusive rank of the threads with `flag` set; returns the block total in `total`
__device__ __forceinline__ unsigned int block_rank(bool flag, unsigned int* warp_sums, unsigned int& total) {
    const unsigned int lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    unsigned int m = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) warp_sums[warp] = (unsigned int)__popc(m);
    __syncthreads();
    unsigned int before = 0, all = 0;
#pragma unroll
    for (unsigned int w = 0; w < WF_SHADE_THREADS / 32; ++w) {
        unsigned int v = warp_sums[w];
        before += w < warp ? v : 0u;
        all += v;
    }
    __syncthreads();
    total = all;
    return before + (unsigned int)__popc(m & ((1u << lane) - 1u));
}

// Shade: one thread per queued path; blocks are class-uniform.  Two atomics per BLOCK: one reserves camera-path
// numbers for the regenerated paths, one reserves the block's range of the next active queue.
__global__ void __launch_bounds__(WF_SHADE_THREADS, 3) wf_shade_kernel(DSceneView S, WfView W, int parity) {
    __shared__ unsigned int warp_sums[WF_SHADE_THREADS / 32];
    __shared__ unsigned long long path_base;
    __shared__ unsigned int queue_base;
    WfCounters* ctr = W.ctr;
    const DRenderParams P = W.job->P;
    AccumFx* __restrict__ accum = W.job->accum;
    int cls = -1;
    unsigned int item = 0, n_in_class = 0;
    {
        unsigned int b = blockIdx.x, start = 0;
#pragma unroll
        for (int c = 0; c < WF_CLASSES; ++c) {
            unsigned int n = ctr->class_count[c];
            unsigned int nb = (n + WF_SHADE_THREADS - 1) / WF_SHADE_THREADS;
            if (cls < 0 && b < start + nb) {
                cls = c;
                item = (b - start) * WF_SHADE_THREADS + threadIdx.x;
                n_in_class = n;
            }
            start += nb;
        }
    }
    if (cls >= 0) {  // uniform per block
        unsigned int* __restrict__ Qout = W.q[parity ^ 1];
        bool valid = item < n_in_class;
        unsigned int slot = 0;
        WfSlot s;
        bool alive = false, dead = false;
        V3 radiance = v3(0.f, 0.f, 0.f);
        if (valid) {
            slot = W.cq[(size_t)cls * W.n_slots + item];
            const float4* rec = reinterpret_cast<const float4*>(W.slots + slot);
            s.A = ld_stream(rec), s.B = ld_stream(rec + 1), s.C = ld_stream(rec + 2), s.D = ld_stream(rec + 3);
            alive = wf_shade(S, P, s, radiance);
            dead = !alive;
        }
        // terminated paths: deposit, then regenerate in place while camera paths remain
        unsigned int n_dead;
        unsigned int dead_rank = block_rank(dead, warp_sums, n_dead);
        if (n_dead) {
            if (threadIdx.x == 0) path_base = atomicAdd(&ctr->next_path, (unsigned long long)n_dead);
            __syncthreads();
            if (dead) {
                unsigned int pixel = __float_as_uint(s.A.w);
                AccumFx* dst = accum + 3 * (size_t)pixel;
                if (radiance.x != 0.f) atomicAdd(dst + 0, radiance_fixed(radiance.x));
                if (radiance.y != 0.f) atomicAdd(dst + 1, radiance_fixed(radiance.y));
                if (radiance.z != 0.f) atomicAdd(dst + 2, radiance_fixed(radiance.z));
                unsigned long long path = path_base + dead_rank;
                if (path < ctr->total_paths) {
                    wf_init_path(S, W.job->cam, P, path, s);
                    alive = true;
                }
            }
        }
        // survivors (scattered or regenerated) -> next active queue
        unsigned int n_alive;
        unsigned int alive_rank = block_rank(alive, warp_sums, n_alive);
        if (n_alive) {
            if (threadIdx.x == 0) queue_base = atomicAdd(&ctr->q_count[parity ^ 1], n_alive);
            __syncthreads();
            if (alive) {
                float4* rec = reinterpret_cast<float4*>(W.slots + slot);
                st_stream(rec, s.A), st_stream(rec + 1, s.B), st_stream(rec + 2, s.C), st_stream(rec + 3, s.D);
                Qout[queue_base + alive_rank] = slot;
            }
        }
    }
    // ---- last block out: clear the class queues and the queue this round consumed
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned int ticket = atomicAdd(&ctr->done_shade, 1u);
        if (ticket == gridDim.x - 1) {
            ctr->done_shade = 0;
            for (int c = 0; c <= WF_CLASSES; ++c) ctr->class_count[c] = 0;
            ctr->q_count[parity] = 0;
            __threadfence();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
namespace {

template <class T>
int wf_alloc(WavefrontState* w, T** out, size_t count) {
    void* p = nullptr;
    CU_TRY(cudaMalloc(&p, count * sizeof(T)));
    w->owned.push_back(p);
    *out = (T*)p;
    return RT_OK;
}

int ensure_state(RtScene* s, unsigned int n_slots) {
    if (s->wf && s->wf->view.n_slots == n_slots) return RT_OK;
    free_wavefront(s);
    WavefrontState* w = new WavefrontState();
    s->wf = w;
    int rc;
    if ((rc = wf_alloc(w, &w->view.slots, n_slots)) || (rc = wf_alloc(w, &w->view.q[0], n_slots)) || (rc = wf_alloc(w, &w->view.q[1], n_slots)) ||
        (rc = wf_alloc(w, &w->view.cq, (size_t)n_slots * WF_CLASSES)) || (rc = wf_alloc(w, &w->view.ctr, 1)) || (rc = wf_alloc(w, &w->view.job, 1)))
        return rc;
    w->view.n_slots = n_slots;
    CU_TRY(cudaMallocHost(&w->h_snapshot, sizeof(WfCounters)));
    CU_TRY(cudaEventCreateWithFlags(&w->snap_event, cudaEventDisableTiming));
    int per_sm = 0, sms = 0;
    CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wf_extend_kernel, WF_EXT_THREADS, 0));
    CU_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device));
    w->ext_blocks = std::max(1, per_sm) * std::max(1, sms);  // persistent: exactly what is co-resident
    return RT_OK;
}

void launch_round(const RtScene* s, WavefrontState* w, int parity, cudaStream_t stream) {
    unsigned int shade_blocks = (w->view.n_slots + WF_SHADE_THREADS - 1) / WF_SHADE_THREADS + WF_CLASSES;
    wf_extend_kernel<<<w->ext_blocks, WF_EXT_THREADS, 0, stream>>>(s->view, w->view, parity);
    wf_shade_kernel<<<shade_blocks, WF_SHADE_THREADS, 0, stream>>>(s->view, w->view, parity);
}

}  // namespace

void free_wavefront(RtScene* s) {
    WavefrontState* w = s->wf;
    if (!w) return;
    if (w->graph) cudaGraphExecDestroy(w->graph);
    for (void* p : w->owned) cudaFree(p);
    if (w->h_snapshot) cudaFreeHost(w->h_snapshot);
    if (w->snap_event) cudaEventDestroy(w->snap_event);
    delete w;
    s->wf = nullptr;
}

int launch_wavefront(RtScene* s, const DCamera& cam, const RtParams* p, int begin, int count, AccumFx* d_accum, cudaStream_t stream, RtProgressFn cb,
                     void* user, int* launches) {
    if (p->max_depth > WF_DEPTH_MASK) return set_error(RT_ERR_UNSUPPORTED, "wavefront pipeline: max_depth above %d", WF_DEPTH_MASK);
    if (p->max_depth <= 0) return RT_OK;  // every path returns Color::ZERO at once (raytrace.rs:87-89)
    if (s->flat.prims.size() >= (1u << 24)) return set_error(RT_ERR_UNSUPPORTED, "wavefront pipeline: more than 2^24 primitives");
    unsigned long long total = (unsigned long long)p->width * p->height * (unsigned long long)count;
    // 4 Mi paths in flight (256 MB pool): measured on C4 with the final device code — 0.5 Mi 714, 1 Mi 961, 2 Mi 1087, 4 Mi 1173,
    // 8 Mi 1106, 16 Mi 1056 Mpaths/s.  Per-round launch gaps and tails outweigh L2 residency of the pool.
    unsigned int n_slots = 1u << 22;
    if (const char* e = getenv("RT_WF_SLOTS")) n_slots = std::max(1024u, (unsigned int)strtoul(e, nullptr, 10));
    int rc = ensure_state(s, n_slots);
    if (rc != RT_OK) return rc;
    WavefrontState* w = s->wf;
    DRenderParams P = device_params(p, begin, 1, 1);
    unsigned int n_init = (unsigned int)std::min<unsigned long long>(total, n_slots);

    wf_reset_kernel<<<1, 1, 0, stream>>>(w->view, n_init, total, cam, P, d_accum);
    wf_init_kernel<<<(n_init + 255) / 256, 256, 0, stream>>>(s->view, w->view, cam, P, n_init);
    CU_TRY(cudaGetLastError());
    *launches += 2;

    // a batch of rounds as one CUDA graph (the kernels' parameters only depend on the round's parity)
    const int kRoundsPerBatch = 32;  // even: every batch starts at parity 0
    if (!w->graph) {
        cudaStream_t cap;
        CU_TRY(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
        cudaGraph_t g = nullptr;
        CU_TRY(cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
        for (int k = 0; k < kRoundsPerBatch; ++k) launch_round(s, w, k & 1, cap);
        CU_TRY(cudaStreamEndCapture(cap, &g));
        CU_TRY(cudaGraphInstantiate(&w->graph, g, 0));
        cudaGraphDestroy(g);
        cudaStreamDestroy(cap);
    }

    // enqueue batches; look at the counters one batch behind
    bool snapshot_pending = false;
    for (int batch = 0;; ++batch) {
        CU_TRY(cudaGraphLaunch(w->graph, stream));
        *launches += 2 * kRoundsPerBatch;
        if (snapshot_pending) {
            CU_TRY(cudaEventSynchronize(w->snap_event));
            const WfCounters& c = *w->h_snapshot;
            if (cb) cb((int)std::min<unsigned long long>(c.next_path / ((unsigned long long)p->width * p->height), (unsigned long long)count), count, user);
            if (c.next_path >= c.total_paths && c.q_count[0] == 0) break;  // the batch after the snapshot ran on an empty pool
        }
        CU_TRY(cudaMemcpyAsync(w->h_snapshot, w->view.ctr, sizeof(WfCounters), cudaMemcpyDeviceToHost, stream));
        CU_TRY(cudaEventRecord(w->snap_event, stream));
        snapshot_pending = true;
        if (batch > (1 << 22)) return set_error(RT_ERR_CUDA, "wavefront pipeline did not terminate");
    }
    // rays of this call -> the scene's ray counter (same place the megakernel accumulates into)
    CU_TRY(cudaMemcpyAsync(w->h_snapshot, w->view.ctr, sizeof(WfCounters), cudaMemcpyDeviceToHost, stream));
    CU_TRY(cudaStreamSynchronize(stream));
    CU_TRY(cudaMemcpyAsync(s->d_rays, &w->h_snapshot->rays, sizeof(unsigned long long), cudaMemcpyHostToDevice, stream));
    return RT_OK;
}

}  // namespace rtb
