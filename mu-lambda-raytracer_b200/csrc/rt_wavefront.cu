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
    float* accum;
};

struct WfView {
    float4 *A, *B, *C, *D;  // path pool, n_slots each
    unsigned int* q[2];     // active queues
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
#define WF_SHADE_THREADS 256
#define WF_DONE ((int)0x80000000)
#define WF_REFILL 8  // lanes that must be idle before a warp stops traversing to retire / fetch rays

// ---------------------------------------------------------------------------------------------------------------
__global__ void wf_reset_kernel(WfView W, unsigned int n_init, unsigned long long total, DCamera cam, DRenderParams P, float* accum) {
    W.job->cam = cam, W.job->P = P, W.job->accum = accum;
    WfCounters* c = W.ctr;
    c->q_count[0] = n_init, c->q_count[1] = 0, c->q_head = 0;
    for (int k = 0; k <= WF_CLASSES; ++k) c->class_count[k] = 0;
    c->done_extend = c->done_shade = 0, c->rounds = 0;
    c->next_path = n_init, c->total_paths = total, c->rays = 0;
}

__global__ void wf_init_kernel(WfView W, DCamera cam, DRenderParams P, unsigned int n_init) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_init) return;
    WfSlot s;
    wf_init_path(cam, P, i, s);
    W.A[i] = s.A, W.B[i] = s.B, W.C[i] = s.C, W.D[i] = s.D;
    W.q[0][i] = i;
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(WF_EXT_THREADS) wf_extend_kernel(DSceneView S, WfView W, int parity) {
    WfCounters* ctr = W.ctr;
    const DRenderParams P = W.job->P;
    const unsigned int qn = ctr->q_count[parity];
    const unsigned int* __restrict__ Q = W.q[parity];
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int lt_mask = (1u << lane) - 1u;

    int stack[RTB_BVH_STACK];
    int sp = 0;
    int cur = WF_DONE;
    bool has = false;
    bool exhausted = qn == 0;
    unsigned int slot = 0, pixel = 0, sample = 0, flags = 0;
    Ray r;
    r.o = r.d = v3(0.f, 0.f, 0.f);
    V3 inv = v3(0.f, 0.f, 0.f);
    float t_best = RTB_INF;
    int prim_best = -1, face_best = 0, origin_prim = -1;
    const int root_link = S.n_prims > 0 ? (int)as_uint(ld4(S.nodes).w) : WF_DONE;

    for (;;) {
        // ---- retire finished rays in batches: media sample, hit record, class queue
        bool fin = has && cur == WF_DONE;
        unsigned int fin_mask = __ballot_sync(0xffffffffu, fin);
        unsigned int trav_mask = __ballot_sync(0xffffffffu, has && cur != WF_DONE);
        unsigned int empty_mask = __ballot_sync(0xffffffffu, !has);
        if (fin_mask && (__popc(fin_mask | empty_mask) >= WF_REFILL || trav_mask == 0u)) {
            int cls = -1;
            if (fin) {
                float t_out;
                int code;
                cls = wf_finish_extend(S, P, r, pixel, sample, flags, t_best, prim_best, face_best, t_out, code);
                float2 rec;
                rec.x = t_out, rec.y = __int_as_float(code);
                *reinterpret_cast<float2*>(&W.D[slot]) = rec;
            }
#pragma unroll
            for (int c = 0; c < WF_CLASSES; ++c) {
                unsigned int m = __ballot_sync(0xffffffffu, cls == c);
                if (m) {
                    unsigned int base = 0;
                    int leader = __ffs(m) - 1;
                    if ((int)lane == leader) base = atomicAdd(&ctr->class_count[c], (unsigned int)__popc(m));
                    base = __shfl_sync(0xffffffffu, base, leader);
                    if (cls == c) W.cq[(size_t)c * W.n_slots + base + __popc(m & lt_mask)] = slot;
                }
            }
            if (fin) has = false;
            empty_mask |= fin_mask;
        }
        // ---- refill empty lanes from the active queue: one atomic per warp
        if (!exhausted && empty_mask && (__popc(empty_mask) >= WF_REFILL || trav_mask == 0u)) {
            unsigned int n = (unsigned int)__popc(empty_mask);
            unsigned int base = 0;
            if (lane == 0) base = atomicAdd(&ctr->q_head, n);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (!has) {
                unsigned int my = base + (unsigned int)__popc(empty_mask & lt_mask);
                if (my < qn) {
                    slot = Q[my];
                    float4 a = W.A[slot], b = W.B[slot], d = W.D[slot];
                    r.o = v3(a.x, a.y, a.z), r.d = v3(b.x, b.y, b.z);
                    pixel = __float_as_uint(a.w), flags = __float_as_uint(b.w), sample = __float_as_uint(d.w);
                    origin_prim = (int)__float_as_uint(d.z);
                    inv = v3(1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z);
                    t_best = RTB_INF, prim_best = -1, face_best = 0;
                    sp = 0, cur = root_link;
                    has = true;
                }
            }
            if (base + n >= qn) exhausted = true;
        }
        if (__ballot_sync(0xffffffffu, has) == 0u) break;  // nothing in flight and the queue is empty

        // ---- traversal: the warp votes between "one inner-node step" and "one leaf-primitive step"
        for (;;) {
            bool inner = has && cur >= 0;
            bool leaf = has && cur < 0 && cur != WF_DONE;
            unsigned int mi = __ballot_sync(0xffffffffu, inner), ml = __ballot_sync(0xffffffffu, leaf);
            if ((mi | ml) == 0u) break;
            if (__popc(mi) >= __popc(ml)) {
                if (inner) {
                    const char* base = reinterpret_cast<const char*>(S.nodes + cur);
                    float4 l0 = ld4(base), l1 = ld4(base + 16), r0 = ld4(base + 32), r1 = ld4(base + 48);
                    float tl, tr;
                    bool hl = slab_node(l0, l1, r.o, inv, RTB_T_MIN, t_best, tl);
                    bool hr = slab_node(r0, r1, r.o, inv, RTB_T_MIN, t_best, tr);
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
                        cur = sp > 0 ? stack[--sp] : WF_DONE;
                    }
                }
            } else {
                if (leaf) {
                    int v = ~cur;
                    int first = v & 0xFFFFFF, count = v >> 24;
                    PrimRec p = load_prim(S.prims + first);
                    float t;
                    int face;
                    if (hit_prim(S, p, r, RTB_T_MIN, t_best, first == origin_prim, (int)((flags >> WF_FACE_SHIFT) & 7u), t, face))
                        t_best = t, prim_best = first, face_best = face;
                    if (count > 1) cur = ~((first + 1) | ((count - 1) << 24));
                    else cur = sp > 0 ? stack[--sp] : WF_DONE;
                }
            }
            if (!exhausted) {
                unsigned int idle = __ballot_sync(0xffffffffu, !has || cur == WF_DONE);
                if (__popc(idle) >= WF_REFILL) break;
            }
        }
    }

    // ---- last block out: account the rays of this round and rewind the cursors the shade stage does not own
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
__global__ void __launch_bounds__(WF_SHADE_THREADS) wf_shade_kernel(DSceneView S, WfView W, int parity) {
    WfCounters* ctr = W.ctr;
    const DRenderParams P = W.job->P;
    float* __restrict__ accum = W.job->accum;
    // blocks are class-uniform: block b belongs to the class whose block range contains it
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
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int lt_mask = (1u << lane) - 1u;
    unsigned int* __restrict__ Qout = W.q[parity ^ 1];
    bool valid = cls >= 0 && item < n_in_class;
    if (cls >= 0) {  // uniform per block
        unsigned int slot = 0;
        WfSlot s;
        bool alive = false, dead = false;
        V3 radiance = v3(0.f, 0.f, 0.f);
        if (valid) {
            slot = W.cq[(size_t)cls * W.n_slots + item];
            s.A = W.A[slot], s.B = W.B[slot], s.C = W.C[slot], s.D = W.D[slot];
            alive = wf_shade(S, P, s, radiance);
            dead = !alive;
        }
        // terminated paths: deposit, then regenerate in place while camera paths remain
        unsigned int dead_mask = __ballot_sync(0xffffffffu, dead);
        if (dead_mask) {
            unsigned long long base = 0;
            int leader = __ffs(dead_mask) - 1;
            if ((int)lane == leader) base = atomicAdd(&ctr->next_path, (unsigned long long)__popc(dead_mask));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (dead) {
                unsigned int pixel = __float_as_uint(s.A.w);
                float* dst = accum + 3 * (size_t)pixel;
                if (radiance.x != 0.f) atomicAdd(dst + 0, radiance.x);
                if (radiance.y != 0.f) atomicAdd(dst + 1, radiance.y);
                if (radiance.z != 0.f) atomicAdd(dst + 2, radiance.z);
                unsigned long long path = base + (unsigned long long)__popc(dead_mask & lt_mask);
                if (path < ctr->total_paths) {
                    wf_init_path(W.job->cam, P, path, s);
                    alive = true;
                }
            }
        }
        // survivors (scattered or regenerated) -> next active queue, compacted with ballot ranks
        unsigned int alive_mask = __ballot_sync(0xffffffffu, alive);
        if (alive_mask) {
            unsigned int base = 0;
            int leader = __ffs(alive_mask) - 1;
            if ((int)lane == leader) base = atomicAdd(&ctr->q_count[parity ^ 1], (unsigned int)__popc(alive_mask));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (alive) {
                W.A[slot] = s.A, W.B[slot] = s.B, W.C[slot] = s.C, W.D[slot] = s.D;
                Qout[base + __popc(alive_mask & lt_mask)] = slot;
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
    if ((rc = wf_alloc(w, &w->view.A, n_slots)) || (rc = wf_alloc(w, &w->view.B, n_slots)) || (rc = wf_alloc(w, &w->view.C, n_slots)) ||
        (rc = wf_alloc(w, &w->view.D, n_slots)) || (rc = wf_alloc(w, &w->view.q[0], n_slots)) || (rc = wf_alloc(w, &w->view.q[1], n_slots)) ||
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

int launch_wavefront(RtScene* s, const DCamera& cam, const RtParams* p, int begin, int count, float* d_accum, cudaStream_t stream, RtProgressFn cb,
                     void* user, int* launches) {
    if (p->max_depth > WF_DEPTH_MASK) return set_error(RT_ERR_UNSUPPORTED, "wavefront pipeline: max_depth above %d", WF_DEPTH_MASK);
    if (p->max_depth <= 0) return RT_OK;  // every path returns Color::ZERO at once (raytrace.rs:87-89)
    if (s->flat.prims.size() >= (1u << 24)) return set_error(RT_ERR_UNSUPPORTED, "wavefront pipeline: more than 2^24 primitives");
    unsigned long long total = (unsigned long long)p->width * p->height * (unsigned long long)count;
    unsigned int n_slots = 1u << 20;
    if (const char* e = getenv("RT_WF_SLOTS")) n_slots = std::max(1024u, (unsigned int)strtoul(e, nullptr, 10));
    int rc = ensure_state(s, n_slots);
    if (rc != RT_OK) return rc;
    WavefrontState* w = s->wf;
    DRenderParams P = device_params(p, begin, 1, 1);
    unsigned int n_init = (unsigned int)std::min<unsigned long long>(total, n_slots);

    wf_reset_kernel<<<1, 1, 0, stream>>>(w->view, n_init, total, cam, P, d_accum);
    wf_init_kernel<<<(n_init + 255) / 256, 256, 0, stream>>>(w->view, cam, P, n_init);
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
