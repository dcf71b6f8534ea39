// The opaque RtScene of include/rt_b200.h and helpers shared by the CUDA translation units (rt_api.cu, rt_wavefront.cu).
#pragma once
#include <cuda_runtime.h>

#include <utility>
#include <vector>

#include "dev_cache.h"
#include "flatten.h"
#include "internal.h"
#include "rt_types.h"

#define CU_TRY(expr)                                                                                                \
    do {                                                                                                            \
        cudaError_t e_ = (expr);                                                                                    \
        if (e_ != cudaSuccess) return rtb::set_error(RT_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

namespace rtb {

struct DeviceGuard {  // run on the scene's device, restore the caller's afterwards
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (dev != prev) cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        int cur = -1;
        cudaGetDevice(&cur);
        if (prev >= 0 && cur != prev) cudaSetDevice(prev);
    }
};

struct WavefrontState;  // rt_wavefront.cu
struct PersistState;    // rt_persist.cu

DRenderParams device_params(const RtParams* p, int first_sample, int spi, int chunks);
int persist_layout_used(const RtScene* s, const RtParams* p);
void persist_preload(const RtScene* s);

struct RowProgress {  // see rtb::row_progress (rt_api.cu)
    RtProgressFn cb;
    void* user;
    int rows, reported;
};
void row_progress(int done, int total, void* user);

}  // namespace rtb

struct RtScene {
    int device = 0;
    rtb::DSceneView view{};
    rtb::FlatScene flat;  // host copy (info + hit -> description node mapping)
    std::vector<std::pair<void*, size_t>> owned;  // (block, bytes): returned to the device cache on destruction
    std::vector<cudaArray_t> arrays;
    std::vector<std::pair<size_t, size_t>> array_extent;
    std::vector<cudaTextureObject_t> textures;
    int64_t device_bytes = 0;
    rtb::OwnedDesc* desc = nullptr;  // deep copy of the description (sub-tree queries of rt_intersect_batch)
    // scratch of rt_render (host-buffer entry point), grown on demand
    rtb::AccumFx* d_accum = nullptr;  // fixed-point radiance sums (rt_types.h)
    float* d_accum_f = nullptr;       // the same as floats, for callers that ask for them
    int32_t* d_rgb = nullptr;
    size_t scratch_values = 0;
    unsigned long long* d_rays = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    rtb::WavefrontState* wf = nullptr;  // global path pool + queues (RT_PIPELINE_WAVEFRONT), allocated on first use
    rtb::PersistState* ps = nullptr;    // counters of the persistent pipeline (RT_PIPELINE_PERSISTENT)
};

namespace rtb {
int lbvh_build(FlatScene& f, float* build_ms);  // rt_lbvh.cu: BVH over f.prim_bounds on the current device
int scene_scratch(RtScene* scene, size_t n_values);  // grow d_accum / d_accum_f / d_rgb to n_values
// samples [begin, begin + count) of every pixel ADDED into the fixed-point buffer d_accum (memory of scene's device)
int accumulate_fixed(const RtScene* scene, const RtCamera* cam, const RtParams* params, AccumFx* d_accum, cudaStream_t stream, RtProgressFn cb,
                     void* user, RtStats* stats, bool sync_for_stats);
// pipelines: samples [begin, begin+count) of every pixel, ADDED into d_accum; *launches counts kernel launches
int launch_megakernel(const RtScene* s, const DCamera& cam, const RtParams* p, int begin, int count, AccumFx* d_accum, cudaStream_t stream,
                      RtProgressFn cb, void* user, int* launches);
int launch_wavefront(RtScene* s, const DCamera& cam, const RtParams* p, int begin, int count, AccumFx* d_accum, cudaStream_t stream,
                     RtProgressFn cb, void* user, int* launches);
void free_wavefront(RtScene* s);
int launch_persist(RtScene* s, const DCamera& cam, const RtParams* p, int begin, int count, AccumFx* d_accum, cudaStream_t stream,
                   RtProgressFn cb, void* user, int* launches);
bool persist_supports(const RtScene* s, const RtParams* p);
void free_persist(RtScene* s);
}  // namespace rtb
