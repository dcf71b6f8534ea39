#include "dev_cache.h"

#include <map>
#include <mutex>
#include <utility>
#include <vector>

namespace rtb {
namespace {

const size_t kMaxCachedBytes = 512ull << 20;  // per device

struct DeviceCache {
    std::multimap<size_t, void*> blocks;                                  // size -> free block
    std::multimap<std::pair<size_t, size_t>, cudaArray_t> arrays;          // (w, h) -> free uchar4 array
    size_t bytes = 0;
};
std::mutex g_mutex;
std::map<int, DeviceCache> g_cache;

int current_device() {
    int d = 0;
    cudaGetDevice(&d);
    return d;
}

}  // namespace

cudaError_t cache_malloc(void** out, size_t bytes) {
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        DeviceCache& c = g_cache[current_device()];
        auto it = c.blocks.find(bytes);
        if (it != c.blocks.end()) {
            *out = it->second;
            c.blocks.erase(it);
            c.bytes -= bytes;
            return cudaSuccess;
        }
    }
    return cudaMalloc(out, bytes);
}

void cache_free(void* p, size_t bytes) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        DeviceCache& c = g_cache[current_device()];
        if (c.bytes + bytes <= kMaxCachedBytes) {
            c.blocks.emplace(bytes, p);
            c.bytes += bytes;
            return;
        }
    }
    cudaFree(p);
}

cudaError_t cache_malloc_array(cudaArray_t* out, const cudaChannelFormatDesc* desc, size_t width, size_t height) {
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        DeviceCache& c = g_cache[current_device()];
        auto it = c.arrays.find(std::make_pair(width, height));
        if (it != c.arrays.end()) {
            *out = it->second;
            c.arrays.erase(it);
            c.bytes -= width * height * 4;
            return cudaSuccess;
        }
    }
    return cudaMallocArray(out, desc, width, height);
}

void cache_free_array(cudaArray_t a, size_t width, size_t height) {
    if (!a) return;
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        DeviceCache& c = g_cache[current_device()];
        if (c.bytes + width * height * 4 <= kMaxCachedBytes) {
            c.arrays.emplace(std::make_pair(width, height), a);
            c.bytes += width * height * 4;
            return;
        }
    }
    cudaFreeArray(a);
}

void cache_release_all() {
    std::lock_guard<std::mutex> lock(g_mutex);
    int prev = current_device();
    for (auto& kv : g_cache) {
        cudaSetDevice(kv.first);
        for (auto& b : kv.second.blocks) cudaFree(b.second);
        for (auto& a : kv.second.arrays) cudaFreeArray(a.second);
        kv.second.blocks.clear(), kv.second.arrays.clear(), kv.second.bytes = 0;
    }
    cudaSetDevice(prev);
}

}  // namespace rtb
