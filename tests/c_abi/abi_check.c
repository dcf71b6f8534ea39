/* A C99 client of include/rt_b200.h: proves the header is plain C and that the host-only entry points link and run
 * from C (what a cgo / Rust FFI / JNI binding sees).  Prints the canonical hash of a world; no GPU is touched, and
 * without a GPU rt_scene_create must answer RT_ERR_NO_DEVICE (there is no CPU path). */
#include <stdio.h>
#include <string.h>

#include "rt_b200.h"

int main(int argc, char** argv) {
    const char* world = argc > 1 ? argv[1] : "cornell_smoke";
    RtWorldInfo info;
    RtSceneDesc* desc = NULL;
    uint64_t draws = 0;
    uint8_t hash[32];
    int i;
    if (rt_abi_version() != RT_B200_ABI_VERSION) return 2;
    if (rt_world_info(world, &info) != RT_OK) {
        fprintf(stderr, "%s\n", rt_last_error());
        return 3;
    }
    if (info.needs_earthmap) return 4; /* keep the check self-contained */
    if (rt_world_build(world, 42, NULL, 0, 0, &desc, &draws) != RT_OK) return 5;
    if (rt_scene_hash(desc, hash) != RT_OK) return 6;
    for (i = 0; i < 32; ++i) printf("%02x", hash[i]);
    printf(" %d %llu %d\n", desc->n_nodes, (unsigned long long)draws, rt_device_count());
    if (rt_device_count() == 0) {
        RtScene* scene = NULL;
        if (rt_scene_create(desc, 0, &scene) != RT_ERR_NO_DEVICE || scene != NULL) return 7;
    }
    rt_scene_desc_free(desc);
    return 0;
}
