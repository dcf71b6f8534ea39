// Flattened scene layout shared by the host flattener (flatten.cpp) and the device code (rt_device.cuh).
//
// Everything the render loop touches per ray is a 32-byte record fetched as two 128-bit loads:
//   DNode  — one BVH node  (replaces the boxed `enum Node` tree of src/bhv.rs:103-106)
//   DPrim  — one primitive (replaces the Sphere / XYRect / XZRect / YZRect / Block trait objects of
//            src/shapes.rs, with Translate/Rotate of src/transforms.rs folded into an instance record)
// The whole C4 scene is ~1.4 k primitives + ~1 k nodes ≈ 80 KB, i.e. L1/L2 resident; HBM only ever sees the
// accumulation buffer.
#pragma once
#include <stdint.h>

namespace rtb {

// ---- BVH node, 32 B.  Children of an inner node are adjacent (left = a, right = a + 1) so one 64-byte
// fetch brings both child boxes; the box of a node lives in the node itself.
// `a` is the traversal link: >= 0 = index of the left child of an inner node, < 0 = ~(first | count << 24) of a leaf.
struct alignas(16) DNode {
    float lo[3];
    int32_t a;  // traversal link (see above)
    float hi[3];
    int32_t b;  // inner: 0;  leaf: primitive count (informational)
};

// ---- 4-wide BVH node, 128 B (one cache line): the binary tree collapsed so that a ray takes half as many dependent
// fetch steps and every step tests four independent boxes.  Child boxes are stored as CENTRE and HALF-EXTENT:
// child c: ch[c] = (centre.x, centre.y, half.x, half.y), cz[c], hz[c].  With t_c = c * (1/d) - o/d the slab interval of an
// axis is t_c -+ h * |1/d|: three multiply-adds on the FMA pipe instead of two multiply-adds and a min/max pair on the
// ALU pipe, which is the busiest pipe of the traversal (60 % in the round-2 capture).
// link[c]: 16-bit child link in the low half — bit 15 clear = index of an inner DNode4, bit 15 set = ONE primitive
// (index in the low 15 bits) — or RTB_LINK4_EMPTY (all ones) for an unused slot.  The upper half of a used link is
// zero, so that (bits(t_entry) & 0xFFFF0000) | link is directly the 4-byte sort/stack key of rt_device.cuh.
struct alignas(16) DNode4 {
    float ch[4][4];
    float cz[4];
    float hz[4];
    uint32_t link[4];
    uint32_t pad[4];
};
#define RTB_LINK4_EMPTY 0xFFFFFFFFu
#define RTB_LINK4_LEAF 0x8000u
#define RTB_LINK4_DONE 0xFFFFu   // cursor value "traversal finished" (leaf bit set, primitive 0x7FFF is never used)
#define RTB_WIDE_MAX_NODES 32768
#define RTB_WIDE_MAX_PRIMS 32767

enum { PRIM_SPHERE = 0, PRIM_BOX = 1 };
enum {
    PRIM_KIND_MASK = 0x3,
    PRIM_BIG = 0x4,            // sphere evaluated in f64 (|r| >= BIG_SPHERE_RADIUS); v[4] = index into big[]
    PRIM_MOVING = 0x8,         // EXTENSION (RT_NODE_MOVING_SPHERE): centre = v[0..2] + time * moving[v[4]].xyz; never PRIM_BIG
    PRIM_INST_SHIFT = 4,       // bits 4..15: instance index + 1 (0 = world space)
    PRIM_INST_MASK = 0xFFF,
    PRIM_RECT_SHIFT = 16,      // bits 16..17: plane axis + 1 of an XY/XZ/YZ rect stored as a flat box (0 = Block)
    PRIM_RECT_MASK = 0x3,
};

// ---- primitive, 32 B
//   sphere: v = cx, cy, cz, r (signed), [big index bits], 0      (centre already in world space)
//   box:    v = min.xyz, max.xyz in OBJECT space (rects are boxes with min == max on the plane axis)
struct alignas(16) DPrim {
    float v[6];
    uint32_t meta;
    int32_t mat;  // index into materials (== index in the description)
};

struct DBigSphere {
    double c[3];
    double r;
};

// ---- rigid instance: the composed Translate/Rotate chain above a primitive (transforms.rs).
//   rot/trans: object -> world (p_w = rot * p_o + trans).  m1/m2 reproduce how the reference re-applies
//   face-forwarding at the two outermost wrappers (transforms.rs:38,138): the final normal is the sign of
//   the geometric normal n with dot(n, m1*d) < 0, and front_face tells whether the sign chosen one level
//   further in (dot(n, m2*d) < 0) already agreed with it.
struct alignas(16) DInstance {
    float rot[9];
    float trans[3];
    float m1[9];
    float m2[9];
    float pad[2];
};

enum { MAT_LAMBERTIAN = 1, MAT_METAL = 2, MAT_DIELECTRIC = 3, MAT_DIFFUSE_LIGHT = 4, MAT_ISOTROPIC = 5 };
struct alignas(16) DMaterial {
    int32_t kind;
    int32_t tex;  // -1: solid colour folded into albedo
    float fuzz;
    float ior;
    float albedo[3];
    int32_t pad;
};

enum { TEX_SOLID = 1, TEX_CHECKER = 2, TEX_NOISE = 3, TEX_IMAGE = 4 };
struct alignas(16) DTexture {
    int32_t kind;
    int32_t a, b;  // CHECKER: odd/even texture; NOISE: perlin table; IMAGE: image
    float scale;
    float color[3];
    int32_t needs_uv;  // an image texture is reachable from here (through checkers)
};
#define RTB_CHECKER_DEPTH 8

// ---- constant-density medium (volumes.rs:7-65): boundary + isotropic phase material.  The boundary is any Hittable in
// the reference: media_prims[first .. first + count) are its surface primitives; `boundary` repeats the first one so
// that the usual single-primitive boundary costs no extra fetch.
struct alignas(16) DMedium {
    DPrim boundary;
    float neg_inv_density;
    int32_t mat;
    int32_t first;
    int32_t count;
};

// ---- accumulation: radiance sums are 64-bit FIXED-POINT integers (32 fractional bits, RT_ACCUM_FIXED_ONE of the ABI).
// Integer addition is associative, so the sums — and with them the image — do not depend on the order in which the
// device's reductions land, nor on how the samples are split over launches or GPUs: one seed gives one image, bit for
// bit, as the reference's per-row PCG streams do (src/raytrace.rs:179,197).
typedef unsigned long long AccumFx;

struct DCamera {  // camera.rs:3-12, computed on the host in f64 by Camera::new's formulas
    float origin[3];
    float lower_left[3];  // lower_left_corner - origin
    float horizontal[3];
    float vertical[3];
    float u[3];
    float v[3];
    float lens_radius;
    float time0, time1;  // EXTENSION: shutter interval of the ray time (0, 0 = none)
};

#define RTB_PERLIN_POINTS 1024
#define RTB_MAX_MEDIA 4096  // free-flight uniforms come four per Philox block (rt_device.cuh: sample_media)
#define RTB_BVH_STACK 48
#define RTB_WIDE_STACK 64   // 4-byte keys of the 4-wide traversal (flatten.cpp falls back to the binary tree beyond it)
#define RTB_BIG_SPHERE_RADIUS 100.0

struct DImage {
    unsigned long long tex;  // cudaTextureObject_t (device) / const uint8_t* RGBA8 (host emulation)
    int32_t width, height;
};

// ---- scene features.  A device function asks SV::feat (SV = the type of the scene view it was handed) before its run-time
// test, so that a kernel instantiated for SceneViewF<mask> carries no code for what the scene does not contain.  This is
// about instruction FOOTPRINT, not about the tests: the persistent kernel executes ~31 KB of instructions per frame on
// final_scene against a 32 KB instruction cache, and code that a scene never runs sits between the lines it does run.
enum : uint32_t {
    F_BOX = 1u,          // box / rect primitives
    F_INSTBOX = 2u,      // ... under a Translate / Rotate chain (tested in object space)
    F_INSTANCE = 4u,     // any primitive under a Translate / Rotate chain (normals re-face-forwarded per wrapper)
    F_MOVING = 8u,       // EXTENSION: moving spheres
    F_BIG = 16u,         // spheres of |r| >= RTB_BIG_SPHERE_RADIUS (f64 roots next to their surface)
    F_MEDIA = 32u,       // constant media
    F_BOXMEDIA = 64u,    // ... with a box boundary
    F_CHECKER = 128u,
    F_NOISE = 256u,
    F_IMAGE = 512u,
    F_SPHERE = 1024u,    // sphere primitives
    F_ALL = 0xFFFFFFFFu,
};

// what every kernel receives by value
struct DSceneView {
    static constexpr uint32_t feat = F_ALL;
    const DNode* nodes;
    const DNode4* nodes4;  // 4-wide collapse of `nodes` (nullptr when the scene exceeds the 16-bit links)
    const DPrim* prims;
    const DBigSphere* big;
    const DInstance* inst;
    const DMaterial* mats;
    const DTexture* texs;
    const DMedium* media;
    const DPrim* media_prims;  // boundary primitives of all media
    const float* perlin_vec;          // n_perlin x 1024 x 4 floats (xyz, pad)
    const unsigned short* perlin_perm;  // n_perlin x 3 x 1024
    const DImage* images;
    const float* moving;  // EXTENSION: n x 4 floats, c1 - c0 of every moving sphere
    int32_t n_nodes, n_nodes4, n_prims, n_media, n_perlin;
    int32_t media_general;  // more than four media, or a boundary of several primitives: the out-of-line sampler
    uint32_t clear_media;   // bit m: nothing but medium m inside its (convex, single-primitive) boundary — see wf_chain_step
    int32_t bg_kind;
    float bg_top[3];
    float bg_bottom[3];
};

template <uint32_t F>
struct SceneViewF : DSceneView {
    static constexpr uint32_t feat = F;
};

struct DRenderParams {
    int32_t width, height;
    int32_t max_depth;
    int32_t sample_begin;     // first sample index of this launch
    int32_t samples_per_item; // consecutive samples integrated by one thread
    int32_t items_per_pixel;  // number of sample chunks in this launch
    int32_t tiles_x, tiles_y; // 8x4 pixel tiles
    uint32_t seed_lo, seed_hi;
    float inv_wm1, inv_hm1;   // 1/(W-1), 1/(H-1)  (raytrace.rs:191-192)
    double inv_npix;          // 1/(W*H): path number -> (pixel, sample) without a 64-bit integer division (rt_persist.cu)
};

}  // namespace rtb
