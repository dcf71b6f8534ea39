"""Host logic of the product (no GPU): the world builder's descriptions hash bit-exact against the oracle's, the
canonical hash is what the header says it is, and librt_b200.so exports exactly the symbols include/rt_b200.h
declares and fails loudly without a device."""
import ctypes as C
import hashlib
import json
import os
import re
import struct

import numpy as np
import pytest

import mu_lambda_raytracer_b200 as rt
from mu_lambda_raytracer_b200 import abi
import support as S

WORLDS = ["simple", "random", "random_chk", "two_spheres", "simple_light", "cornell_box", "cornell_smoke", "earth",
          "debug_perlin", "final_scene"]


def test_world_registry_matches_reference_order():
    assert [w.name() for w in rt.worlds()] == WORLDS  # worlds.rs:471-484
    w = rt.World("final_scene")
    assert w.camera() == {"lookfrom": (478.0, 278.0, -600.0), "lookat": (278.0, 278.0, 0.0), "field_of_view": 40.0}
    assert isinstance(w.background(), rt.BlackBackground) and isinstance(rt.World("random").background(), rt.GradientBackground)
    with pytest.raises(abi.RtError):
        rt.World("no_such_world")


@pytest.mark.parametrize("name", WORLDS)
@pytest.mark.parametrize("seed", [42, 7])
def test_scene_hash_bit_exact_against_oracle(name, seed):
    desc = rt.World(name).build(seed)
    ow = S.OracleWorld(name, seed)
    assert desc.n_draws == ow.draws
    assert desc.hash() == rt.SceneDescription(ow.desc, owned=False).hash()


def test_golden_scene_hashes():
    with open(os.path.join(S.GOLDEN, "scene_hashes.json")) as f:
        want = json.load(f)
    for name in WORLDS:
        assert rt.World(name).build(42).hash() == want[name], name


def canonical_bytes(d):
    """the serialisation documented in csrc/scene_hash.cpp, re-stated with struct.pack"""
    out = [b"RTB200-SCENE-v1\0", struct.pack("<2i", d.root, d.background_kind), struct.pack("<6d", *d.background_top, *d.background_bottom),
           struct.pack("<6i", d.n_nodes, d.n_children, d.n_materials, d.n_textures, d.n_perlins, d.n_images)]
    for i in range(d.n_nodes):
        n = d.nodes[i]
        out.append(struct.pack("<5i8d", n.kind, n.material, n.first_child, n.child_count, n.axis, *n.f))
    out.append(struct.pack("<%di" % d.n_children, *[d.children[i] for i in range(d.n_children)]))
    for i in range(d.n_materials):
        m = d.materials[i]
        out.append(struct.pack("<2i5d", m.kind, m.texture, *m.albedo, m.fuzz, m.ior))
    for i in range(d.n_textures):
        t = d.textures[i]
        out.append(struct.pack("<3i4d", t.kind, t.a, t.b, *t.color, t.scale))
    for i in range(d.n_perlins):
        p = d.perlins[i]
        out.append(np.ctypeslib.as_array(p.ranvec).astype("<f8").tobytes())
        for perm in (p.perm_x, p.perm_y, p.perm_z):
            out.append(np.ctypeslib.as_array(perm).astype("<i4").tobytes())
    for i in range(d.n_images):
        im = d.images[i]
        out.append(struct.pack("<2i", im.width, im.height))
        out.append(C.string_at(im.rgb, 3 * im.width * im.height))
    return b"".join(out)


@pytest.mark.parametrize("name", ["random", "cornell_smoke", "final_scene"])
def test_scene_hash_is_sha256_of_documented_layout(name):
    desc = rt.World(name).build(42)
    assert desc.hash() == hashlib.sha256(canonical_bytes(desc.desc)).hexdigest()


def test_hash_sensitive_to_one_ulp():
    desc = rt.World("cornell_smoke").build(0)
    h0 = desc.hash()
    desc.desc.nodes[0].f[0] = np.nextafter(desc.desc.nodes[0].f[0], 1e9)
    assert desc.hash() != h0


def test_seed_changes_only_random_worlds():
    assert rt.World("cornell_smoke").build(1).hash() == rt.World("cornell_smoke").build(2).hash()
    assert rt.World("random").build(1).hash() != rt.World("random").build(2).hash()


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(S.ROOT, "include", "rt_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", header))
    assert declared == set(abi.PROTOTYPES), declared ^ set(abi.PROTOTYPES)
    lib = abi.load()
    for name in declared:
        assert hasattr(lib, name), f"librt_b200.so does not export {name}"
    assert lib.rt_abi_version() == 3


def test_struct_sizes_match_header():
    # sizes the C compiler gives the header's structs (LP64): guards the ctypes mirror against drift
    assert C.sizeof(abi.RtNode) == 88 and C.sizeof(abi.RtMaterial) == 48 and C.sizeof(abi.RtTexture) == 48
    assert C.sizeof(abi.RtPerlin) == 1024 * 24 + 3 * 4096 and C.sizeof(abi.RtImage) == 16
    assert C.sizeof(abi.RtCamera) == 120 and C.sizeof(abi.RtParams) == 48 and C.sizeof(abi.RtHit) == 48
    assert C.sizeof(abi.RtStats) == 40 and C.sizeof(abi.RtSceneDesc) == 128


def test_no_device_means_error_not_fallback():
    lib = abi.load()
    if lib.rt_device_count() > 0:
        pytest.skip("a CUDA device is present")
    desc = rt.World("cornell_smoke").build(0)
    h = C.c_void_p()
    assert lib.rt_scene_create(desc.ptr, 0, C.byref(h)) == abi.RT_ERR_NO_DEVICE and not h.value
    assert b"no CPU path" in lib.rt_last_error()
    with pytest.raises(abi.RtError):
        rt.Scene(desc)
    cam = S.make_camera((0, 0, 0), (0, 0, -1), 40, 1.0)
    p = abi.RtParams()
    p.width, p.height, p.samples_per_pixel, p.max_depth = 8, 8, 1, 5
    px, sm = np.zeros(1, np.int32), np.zeros(1, np.int32)
    rays, us = np.zeros(6, np.float32), np.zeros(4, np.float32)
    assert lib.rt_generate_rays(C.byref(cam.c), C.byref(p), px.ctypes.data, sm.ctypes.data, 1, rays.ctypes.data, us.ctypes.data) == abi.RT_ERR_NO_DEVICE
    assert lib.rt_measure_peaks(0, C.byref(abi.RtPeaks())) == abi.RT_ERR_NO_DEVICE
    assert lib.rt_tonemap_fixed_device(1, 1, 1, 1, 0, None) == abi.RT_ERR_NO_DEVICE
    lib.rt_release_cached_memory()  # nothing cached, no device: a no-op


def test_bad_arguments_are_rejected():
    lib = abi.load()
    assert lib.rt_scene_hash(None, None) == abi.RT_ERR_INVALID
    ptr = C.POINTER(abi.RtSceneDesc)()
    assert lib.rt_world_build(b"earth", 1, None, 0, 0, C.byref(ptr), None) == abi.RT_ERR_INVALID  # worlds.rs:180 unwrap()
    assert b"earthmap" in lib.rt_last_error()
    assert lib.rt_world_name(99) is None


def test_to_ppm_framing():  # main.rs:144,175-179
    rgb = np.arange(2 * 3 * 3, dtype=np.int32).reshape(2, 3, 3)
    text = rt.to_ppm(rgb)
    lines = text.split("\n")
    assert lines[:3] == ["P3", "3 2", "255"]
    assert lines[3] == "9 10 11" and lines[6] == "0 1 2" and text.endswith("\n") and len(lines) == 3 + 6 + 1


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/rt_b200.h compiled as C99 by gcc and a C client (tests/c_abi/abi_check.c) linked against librt_b200.so:
    the hash it prints for a world equals the one obtained through ctypes"""
    import subprocess
    inc = os.path.join(S.ROOT, "include")
    libdir = os.path.dirname(abi.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", os.path.join(inc, "rt_b200.h")])
    exe = str(tmp_path / "abi_check")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", inc, os.path.join(S.ROOT, "tests", "c_abi", "abi_check.c"), "-o", exe,
                           "-L", libdir, "-lrt_b200", "-Wl,-rpath," + libdir])
    for world in ("cornell_smoke", "random"):
        out = subprocess.run([exe, world], stdout=subprocess.PIPE, text=True, check=True).stdout.split()
        d = rt.World(world).build(42)
        assert out[0] == d.hash() and int(out[1]) == d.desc.n_nodes and int(out[2]) == d.n_draws
