// rt_scene_hash: canonical SHA-256 of a scene description, and the error-string plumbing of the ABI.
//
// The digest is the "bit-exact against the reference's world" check (BASELINE.json north_star, SURVEY §8c
// level 1): every value is serialised little-endian in construction order with f64 bit patterns, so two
// builders agree on the hash iff they emitted the same objects with the same doubles in the same order.
//
// layout: "RTB200-SCENE-v1\0" | root, background_kind (i32) | top[3], bottom[3] (f64)
//         | n_nodes, n_children, n_materials, n_textures, n_perlins, n_images (i32)
//         | nodes: kind, material, first_child, child_count, axis (i32), f[8] (f64)
//         | children (i32) | materials: kind, texture (i32), albedo[3], fuzz, ior (f64)
//         | textures: kind, a, b (i32), color[3], scale (f64)
//         | perlins: ranvec[1024][3] (f64), perm_x, perm_y, perm_z (i32 x 1024 each)
//         | images: width, height (i32), width*height*3 bytes
#include <cstdio>
#include <cstring>

#include "internal.h"

namespace rtb {

static thread_local char g_error[512] = "";

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof g_error, fmt, ap);
    va_end(ap);
    return code;
}
void clear_error() { g_error[0] = 0; }

static const uint32_t K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

static void compress(uint32_t h[8], const uint8_t block[64]) {
    uint32_t w[64];
    for (int i = 0; i < 16; ++i)
        w[i] = ((uint32_t)block[4 * i] << 24) | ((uint32_t)block[4 * i + 1] << 16) | ((uint32_t)block[4 * i + 2] << 8) | block[4 * i + 3];
    for (int i = 16; i < 64; ++i) {
        uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; ++i) {
        uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = hh + S1 + ch + K[i] + w[i];
        uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        hh = g, g = f, f = e, e = d + t1, d = c, c = b, b = a, a = t1 + t2;
    }
    h[0] += a, h[1] += b, h[2] += c, h[3] += d, h[4] += e, h[5] += f, h[6] += g, h[7] += hh;
}

Sha256::Sha256() {
    static const uint32_t init[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    std::memcpy(h, init, sizeof h);
}
void Sha256::update(const void* data, size_t n) {
    const uint8_t* p = (const uint8_t*)data;
    total += n;
    while (n) {
        size_t take = 64 - fill < n ? 64 - fill : n;
        std::memcpy(buf + fill, p, take);
        fill += (uint32_t)take, p += take, n -= take;
        if (fill == 64) {
            compress(h, buf);
            fill = 0;
        }
    }
}
void Sha256::finish(uint8_t out[32]) {
    uint64_t bits = total * 8;
    uint8_t pad = 0x80;
    update(&pad, 1);
    uint8_t zero = 0;
    while (fill != 56) update(&zero, 1);
    uint8_t len[8];
    for (int i = 0; i < 8; ++i) len[i] = (uint8_t)(bits >> (56 - 8 * i));
    update(len, 8);
    for (int i = 0; i < 8; ++i) {
        out[4 * i] = (uint8_t)(h[i] >> 24), out[4 * i + 1] = (uint8_t)(h[i] >> 16);
        out[4 * i + 2] = (uint8_t)(h[i] >> 8), out[4 * i + 3] = (uint8_t)h[i];
    }
}

namespace {
struct Writer {
    Sha256 sha;
    void i32(int32_t v) {
        uint8_t b[4] = {(uint8_t)v, (uint8_t)(v >> 8), (uint8_t)(v >> 16), (uint8_t)(v >> 24)};
        sha.update(b, 4);
    }
    void f64(double v) {
        uint64_t u;
        std::memcpy(&u, &v, 8);
        uint8_t b[8];
        for (int i = 0; i < 8; ++i) b[i] = (uint8_t)(u >> (8 * i));
        sha.update(b, 8);
    }
};
}  // namespace

}  // namespace rtb

using namespace rtb;

extern "C" {

const char* rt_last_error(void) { return g_error; }
int rt_abi_version(void) { return RT_B200_ABI_VERSION; }

int rt_scene_hash(const RtSceneDesc* d, uint8_t out[32]) {
    if (!d || !out) return set_error(RT_ERR_INVALID, "rt_scene_hash: null argument");
    Writer w;
    w.sha.update("RTB200-SCENE-v1", 16);
    w.i32(d->root), w.i32(d->background_kind);
    for (int i = 0; i < 3; ++i) w.f64(d->background_top[i]);
    for (int i = 0; i < 3; ++i) w.f64(d->background_bottom[i]);
    w.i32(d->n_nodes), w.i32(d->n_children), w.i32(d->n_materials), w.i32(d->n_textures), w.i32(d->n_perlins), w.i32(d->n_images);
    for (int i = 0; i < d->n_nodes; ++i) {
        const RtNode& n = d->nodes[i];
        w.i32(n.kind), w.i32(n.material), w.i32(n.first_child), w.i32(n.child_count), w.i32(n.axis);
        for (int k = 0; k < 8; ++k) w.f64(n.f[k]);
    }
    for (int i = 0; i < d->n_children; ++i) w.i32(d->children[i]);
    for (int i = 0; i < d->n_materials; ++i) {
        const RtMaterial& m = d->materials[i];
        w.i32(m.kind), w.i32(m.texture);
        for (int k = 0; k < 3; ++k) w.f64(m.albedo[k]);
        w.f64(m.fuzz), w.f64(m.ior);
    }
    for (int i = 0; i < d->n_textures; ++i) {
        const RtTexture& t = d->textures[i];
        w.i32(t.kind), w.i32(t.a), w.i32(t.b);
        for (int k = 0; k < 3; ++k) w.f64(t.color[k]);
        w.f64(t.scale);
    }
    for (int i = 0; i < d->n_perlins; ++i) {
        const RtPerlin& p = d->perlins[i];
        for (int k = 0; k < RT_PERLIN_POINTS; ++k)
            for (int c = 0; c < 3; ++c) w.f64(p.ranvec[k][c]);
        for (int k = 0; k < RT_PERLIN_POINTS; ++k) w.i32(p.perm_x[k]);
        for (int k = 0; k < RT_PERLIN_POINTS; ++k) w.i32(p.perm_y[k]);
        for (int k = 0; k < RT_PERLIN_POINTS; ++k) w.i32(p.perm_z[k]);
    }
    for (int i = 0; i < d->n_images; ++i) {
        const RtImage& im = d->images[i];
        w.i32(im.width), w.i32(im.height);
        if (!im.rgb) return set_error(RT_ERR_INVALID, "rt_scene_hash: image %d has no pixels", i);
        w.sha.update(im.rgb, (size_t)3 * im.width * im.height);
    }
    w.sha.finish(out);
    return RT_OK;
}

}  // extern "C"
