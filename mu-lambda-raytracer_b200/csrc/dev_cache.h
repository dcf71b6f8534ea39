// Device-memory cache of librt_b200: blocks and texture arrays freed by one scene are handed to the next one.
//
// Why: a host that renders frame after frame (movie.py drives one rt_main-like render per frame; bench.py's e2e leg
// re-creates the scene every step) would otherwise pay cudaMalloc / cudaFree / cudaFreeArray for the same ~20 blocks
// every time, and on this driver those calls stall at random for tens to hundreds of milliseconds (measured: destroying
// a scene took 1.7 ms typically and 120-550 ms every third time).  Blocks are cached per device by exact size, arrays by
// extent; at most kMaxCachedBytes stay cached per device, the rest is really freed.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace rtb {

cudaError_t cache_malloc(void** out, size_t bytes);   // on the current device
void cache_free(void* p, size_t bytes);               // on the current device
cudaError_t cache_malloc_array(cudaArray_t* out, const cudaChannelFormatDesc* desc, size_t width, size_t height);
void cache_free_array(cudaArray_t a, size_t width, size_t height);
void cache_release_all();  // really free everything cached (all devices)

}  // namespace rtb
