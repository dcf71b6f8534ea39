// Internal helpers shared by the host-side translation units of librt_b200 (not part of the ABI).
#pragma once
#include <cstdarg>
#include <cstdint>
#include <vector>

#include "../../include/rt_b200.h"

namespace rtb {

// records the message returned by rt_last_error() (thread-local) and returns `code`
int set_error(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
void clear_error();

// a scene description that owns its arrays (what rt_world_build returns and what RtScene keeps a copy of)
struct OwnedDesc {
    RtSceneDesc d;  // must stay first: rt_world_build hands out &d
    std::vector<RtNode> nodes;
    std::vector<int32_t> children;
    std::vector<RtMaterial> materials;
    std::vector<RtTexture> textures;
    std::vector<RtPerlin> perlins;
    std::vector<RtImage> images;
    std::vector<std::vector<uint8_t>> pixels;
};
void seal_desc(OwnedDesc& o, int root, int background_kind);  // point d at the vectors
OwnedDesc* clone_desc(const RtSceneDesc* src);                // deep copy
void free_desc(OwnedDesc* o);

// SHA-256 (FIPS 180-4), used by rt_scene_hash
struct Sha256 {
    uint32_t h[8];
    uint8_t buf[64];
    uint64_t total = 0;
    uint32_t fill = 0;
    Sha256();
    void update(const void* data, size_t n);
    void finish(uint8_t out[32]);
};

}  // namespace rtb
