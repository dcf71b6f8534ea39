// GPU-side BVH build for scenes too large for the host's SAH sweep (no counterpart in the reference, whose
// BHV::new is a recursive median split on the host, src/bhv.rs:122-145).  A linear BVH after Karras, "Maximizing
// Parallelism in the Construction of BVHs, Octrees, and k-d Trees" (HPG 2012):
//   1. 30-bit Morton code of every primitive's box centre inside the scene bounds, made unique by the primitive index
//      in the low half of a 64-bit key;
//   2. radix sort of the keys (cub::DeviceRadixSort — a library sort: this is scene set-up, not the render path);
//   3. one thread per inner node finds its key range and split from common-prefix lengths (no recursion, no atomics);
//   4. boxes bottom-up: one thread per leaf climbs, the second arrival at a node (atomic ticket) merges its children;
//   5. the tree is written in the layout the traversal kernels walk: 32-byte DNode records, the two children of a node
//      adjacent (rt_types.h), one primitive per leaf, primitives permuted into leaf order.
// Traversal quality is below the host SAH tree's (no cost model), which is why small scenes keep the host builder; the
// build itself is milliseconds where the host sweep is seconds (tests/test_lbvh.py prints both).
#include <cuda_runtime.h>

#include <cub/cub.cuh>

#include <algorithm>
#include <cfloat>
#include <vector>

#include "scene_internal.h"

namespace rtb {
namespace {

__device__ __forceinline__ unsigned int expand_bits10(unsigned int v) {  // 10 bits -> every third bit
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__global__ void lbvh_keys_kernel(const float* __restrict__ bounds, int n, float3 lo, float3 inv_extent, unsigned long long* __restrict__ keys,
                                 int* __restrict__ index) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* b = bounds + 6 * (size_t)i;
    float cx = 0.5f * (b[0] + b[3]), cy = 0.5f * (b[1] + b[4]), cz = 0.5f * (b[2] + b[5]);
    float x = fminf(fmaxf((cx - lo.x) * inv_extent.x * 1024.0f, 0.0f), 1023.0f);
    float y = fminf(fmaxf((cy - lo.y) * inv_extent.y * 1024.0f, 0.0f), 1023.0f);
    float z = fminf(fmaxf((cz - lo.z) * inv_extent.z * 1024.0f, 0.0f), 1023.0f);
    unsigned int m = expand_bits10((unsigned int)x) * 4u + expand_bits10((unsigned int)y) * 2u + expand_bits10((unsigned int)z);
    keys[i] = ((unsigned long long)m << 32) | (unsigned int)i;  // unique: ties between equal codes are split by index
    index[i] = i;
}

__device__ __forceinline__ int prefix(const unsigned long long* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    return __clzll((long long)(keys[i] ^ keys[j]));
}

// Karras 2012, section 4: range and split of inner node i; children as (index << 1 | is_leaf)
__global__ void lbvh_hierarchy_kernel(const unsigned long long* __restrict__ keys, int n, int2* __restrict__ children, int* __restrict__ parent_inner,
                                      int* __restrict__ parent_leaf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = prefix(keys, n, i, i + 1) - prefix(keys, n, i, i - 1) >= 0 ? 1 : -1;
    int dmin = prefix(keys, n, i, i - d);
    int lmax = 2;
    while (prefix(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (prefix(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = prefix(keys, n, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) / 2;
        if (prefix(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    bool left_leaf = lo == gamma, right_leaf = hi == gamma + 1;
    children[i] = make_int2((gamma << 1) | (left_leaf ? 1 : 0), ((gamma + 1) << 1) | (right_leaf ? 1 : 0));
    if (left_leaf) parent_leaf[gamma] = i;
    else parent_inner[gamma] = i;
    if (right_leaf) parent_leaf[gamma + 1] = i;
    else parent_inner[gamma + 1] = i;
    if (i == 0) parent_inner[0] = -1;
}

struct Box6 {
    float v[6];
};
__device__ __forceinline__ Box6 merge(const Box6& a, const Box6& b) {
    Box6 r;
    for (int k = 0; k < 3; ++k) r.v[k] = fminf(a.v[k], b.v[k]), r.v[3 + k] = fmaxf(a.v[3 + k], b.v[3 + k]);
    return r;
}

// boxes and heights bottom-up: the second thread to arrive at a node merges its two children
__global__ void lbvh_refit_kernel(const float* __restrict__ bounds, const int* __restrict__ sorted_index, int n, const int2* __restrict__ children,
                                  const int* __restrict__ parent_inner, const int* __restrict__ parent_leaf, Box6* __restrict__ inner_box,
                                  int* __restrict__ height, unsigned int* __restrict__ ticket) {
    int leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= n) return;
    int node = parent_leaf[leaf];
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(ticket + node, 1u) == 0u) return;  // first arrival: the sibling subtree is not finished yet
        const int2 c = children[node];
        Box6 b[2];
        int h[2];
        for (int k = 0; k < 2; ++k) {
            const int ref = k == 0 ? c.x : c.y;
            if (ref & 1) {
                const float* src = bounds + 6 * (size_t)sorted_index[ref >> 1];
                for (int q = 0; q < 6; ++q) b[k].v[q] = src[q];
                h[k] = 0;
            } else {
                const volatile float* src = (const volatile float*)(inner_box + (ref >> 1));  // written by another thread before its ticket
                for (int q = 0; q < 6; ++q) b[k].v[q] = src[q];
                h[k] = *(const volatile int*)(height + (ref >> 1));
            }
        }
        inner_box[node] = merge(b[0], b[1]);
        height[node] = 1 + max(h[0], h[1]);
        node = parent_inner[node];
    }
}

// the traversal layout: node 0 = root (box + link to its child pair), inner node i's children at 1 + 2i and 2 + 2i
__global__ void lbvh_emit_kernel(const float* __restrict__ bounds, const int* __restrict__ sorted_index, int n, const int2* __restrict__ children,
                                 const Box6* __restrict__ inner_box, DNode* __restrict__ nodes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int2 c = children[i];
    for (int k = 0; k < 2; ++k) {
        const int ref = k == 0 ? c.x : c.y;
        DNode out;
        if (ref & 1) {
            const float* src = bounds + 6 * (size_t)sorted_index[ref >> 1];
            for (int q = 0; q < 3; ++q) out.lo[q] = src[q], out.hi[q] = src[3 + q];
            out.a = ~((ref >> 1) | (1 << 24)), out.b = 1;  // one primitive: primitives are permuted into leaf order
        } else {
            const Box6 b = inner_box[ref >> 1];
            for (int q = 0; q < 3; ++q) out.lo[q] = b.v[q], out.hi[q] = b.v[3 + q];
            out.a = 1 + 2 * (ref >> 1), out.b = 0;
        }
        nodes[1 + 2 * i + k] = out;
    }
    if (i == 0) {
        const Box6 b = inner_box[0];
        DNode root;
        for (int q = 0; q < 3; ++q) root.lo[q] = b.v[q], root.hi[q] = b.v[3 + q];
        root.a = 1, root.b = 0;
        nodes[0] = root;
    }
}

template <class T>
struct DeviceBuffer {  // freed on every exit
    T* p = nullptr;
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)); }
    ~DeviceBuffer() {
        if (p) cudaFree(p);
    }
};

}  // namespace

// Builds f.nodes on the current device from f.prim_bounds and permutes f.prims / f.prim_node / f.prim_bounds into leaf
// order.  *build_ms = device time of the five steps (CUDA events).
int lbvh_build(FlatScene& f, float* build_ms) {
    const int n = (int)f.prims.size();
    if (n < 2 || f.prims.size() >= (1u << 24)) return set_error(RT_ERR_UNSUPPORTED, "GPU BVH build: needs 2 .. 2^24 - 1 primitives");
    if (f.prim_bounds.size() != 6 * (size_t)n) return set_error(RT_ERR_INVALID, "GPU BVH build: primitive bounds missing");
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) {
            float c = 0.5f * (f.prim_bounds[6 * (size_t)i + k] + f.prim_bounds[6 * (size_t)i + 3 + k]);
            lo[k] = std::min(lo[k], c), hi[k] = std::max(hi[k], c);
        }
    float3 origin = make_float3(lo[0], lo[1], lo[2]);
    float3 inv = make_float3(1.0f / std::max(hi[0] - lo[0], 1e-30f), 1.0f / std::max(hi[1] - lo[1], 1e-30f), 1.0f / std::max(hi[2] - lo[2], 1e-30f));

    DeviceBuffer<float> bounds;
    DeviceBuffer<unsigned long long> keys, keys_sorted;
    DeviceBuffer<int> index, index_sorted, parent_inner, parent_leaf, height;
    DeviceBuffer<int2> children;
    DeviceBuffer<Box6> inner_box;
    DeviceBuffer<unsigned int> ticket;
    DeviceBuffer<DNode> nodes;
    DeviceBuffer<unsigned char> temp;
    CU_TRY(bounds.alloc(6 * (size_t)n));
    CU_TRY(keys.alloc(n));
    CU_TRY(keys_sorted.alloc(n));
    CU_TRY(index.alloc(n));
    CU_TRY(index_sorted.alloc(n));
    CU_TRY(parent_inner.alloc(n));
    CU_TRY(parent_leaf.alloc(n));
    CU_TRY(height.alloc(n));
    CU_TRY(children.alloc(n));
    CU_TRY(inner_box.alloc(n));
    CU_TRY(ticket.alloc(n));
    CU_TRY(nodes.alloc(2 * (size_t)n - 1));
    size_t temp_bytes = 0;
    CU_TRY(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, keys.p, keys_sorted.p, index.p, index_sorted.p, n));
    CU_TRY(temp.alloc(temp_bytes));
    CU_TRY(cudaMemcpy(bounds.p, f.prim_bounds.data(), 6 * (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
    CU_TRY(cudaMemset(ticket.p, 0, (size_t)n * sizeof(unsigned int)));

    cudaEvent_t e0, e1;
    CU_TRY(cudaEventCreate(&e0));
    CU_TRY(cudaEventCreate(&e1));
    const int T = 256, B = (n + T - 1) / T;
    CU_TRY(cudaEventRecord(e0, 0));
    lbvh_keys_kernel<<<B, T>>>(bounds.p, n, origin, inv, keys.p, index.p);
    CU_TRY(cub::DeviceRadixSort::SortPairs(temp.p, temp_bytes, keys.p, keys_sorted.p, index.p, index_sorted.p, n));
    lbvh_hierarchy_kernel<<<B, T>>>(keys_sorted.p, n, children.p, parent_inner.p, parent_leaf.p);
    lbvh_refit_kernel<<<B, T>>>(bounds.p, index_sorted.p, n, children.p, parent_inner.p, parent_leaf.p, inner_box.p, height.p, ticket.p);
    lbvh_emit_kernel<<<B, T>>>(bounds.p, index_sorted.p, n, children.p, inner_box.p, nodes.p);
    CU_TRY(cudaEventRecord(e1, 0));
    CU_TRY(cudaEventSynchronize(e1));
    CU_TRY(cudaGetLastError());
    float ms = 0.f;
    CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0), cudaEventDestroy(e1);
    if (build_ms) *build_ms = ms;

    int depth = 0;
    CU_TRY(cudaMemcpy(&depth, height.p, sizeof(int), cudaMemcpyDeviceToHost));  // height of inner node 0 = the root
    if (depth > RTB_BVH_STACK - 2) return set_error(RT_ERR_UNSUPPORTED, "GPU BVH build: tree depth %d exceeds the traversal stack", depth);
    f.bvh_depth = depth;
    f.nodes.resize(2 * (size_t)n - 1);
    CU_TRY(cudaMemcpy(f.nodes.data(), nodes.p, f.nodes.size() * sizeof(DNode), cudaMemcpyDeviceToHost));
    std::vector<int> order(n);
    CU_TRY(cudaMemcpy(order.data(), index_sorted.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<DPrim> p2(n);
    std::vector<int32_t> n2(n);
    std::vector<float> b2(6 * (size_t)n);
    for (int i = 0; i < n; ++i) {
        p2[i] = f.prims[order[i]], n2[i] = f.prim_node[order[i]];
        std::copy(f.prim_bounds.begin() + 6 * (size_t)order[i], f.prim_bounds.begin() + 6 * (size_t)order[i] + 6, b2.begin() + 6 * (size_t)i);
    }
    f.prims.swap(p2), f.prim_node.swap(n2), f.prim_bounds.swap(b2);
    f.nodes4.clear(), f.bvh4_depth = 0;
    f.built_on_device = true;
    return RT_OK;
}

}  // namespace rtb
