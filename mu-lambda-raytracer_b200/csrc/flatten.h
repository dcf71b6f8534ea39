// Host-side flattening of a scene description into the device layout of rt_types.h.
#pragma once
#include <string>
#include <vector>

#include "../../include/rt_b200.h"
#include "rt_types.h"

namespace rtb {

struct FlatScene {
    std::vector<DNode> nodes;       // BVH over the surface primitives, root = 0
    std::vector<DNode4> nodes4;     // the same tree collapsed 4-wide (empty when it does not fit 16-bit links or leaves hold > 1 primitive)
    std::vector<DPrim> prims;       // in BVH leaf order
    std::vector<int32_t> prim_node; // description node of each primitive
    std::vector<DBigSphere> big;
    std::vector<float> moving;      // 4 floats per moving sphere: c1 - c0 (world space), pad
    std::vector<DInstance> inst;
    std::vector<DMaterial> mats;    // same indexing as the description
    std::vector<DTexture> texs;     // same indexing as the description
    std::vector<DMedium> media;
    std::vector<DPrim> media_prims;   // boundary primitives, medium after medium
    std::vector<int32_t> media_node;  // description node of each medium
    uint32_t features = 0;            // F_* bits of rt_types.h: what the device code has to be able to handle for this scene
    uint32_t clear_media = 0;         // bit m: the interior of medium m's boundary holds no surface (flatten.cpp: find_clear_media)
    std::vector<float> perlin_vec;           // n x 1024 x 4
    std::vector<unsigned short> perlin_perm; // n x 3 x 1024
    struct Image {
        int32_t width, height;
        std::vector<uint8_t> rgba;  // width*height*4, row 0 = top of the file
    };
    std::vector<Image> images;
    int32_t bg_kind = 0;
    float bg_top[3] = {0, 0, 0}, bg_bottom[3] = {0, 0, 0};
    int32_t bvh_depth = 0;
    int32_t bvh4_depth = 0;
    std::vector<float> prim_bounds;    // 6 floats per primitive (world box, rounded outward), in `prims` order
    bool needs_device_build = false;   // BUILD_AUTO left the BVH to the GPU builder (rt_lbvh.cu): `nodes` is still empty
    bool built_on_device = false;
    float device_build_ms = 0.0f;
};

// Flatten the subtree rooted at `root` (normally desc->root).  build: BUILD_NONE leaves the primitives in emission order
// and `nodes` empty (the brute-force test entry point); BUILD_HOST = the SAH sweep + 4-wide collapse on the host;
// BUILD_AUTO = the same for scenes below RTB_GPU_BUILD_MIN primitives (RT_BVH_GPU_MIN overrides), otherwise only
// needs_device_build is set and the caller runs the GPU builder.  Returns RT_OK or an RT_ERR_* code with `err` filled in.
enum { BUILD_NONE = 0, BUILD_HOST = 1, BUILD_AUTO = 2 };
#define RTB_GPU_BUILD_MIN 32768  // above RTB_WIDE_MAX_PRIMS the 4-wide tree does not exist anyway
int flatten_scene(const RtSceneDesc* desc, int32_t root, int build, FlatScene& out, std::string& err);

// Camera::new (camera.rs:15-38) evaluated in f64, narrowed to the device record.
void make_camera(const RtCamera& in, DCamera& out);

}  // namespace rtb
