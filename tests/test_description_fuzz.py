"""The description crosses a C ABI from an untrusted host: whatever it contains, flattening must answer with a status code,
never crash or loop.  Random corruptions of valid descriptions go through the host flattener (the code rt_scene_create
runs before it touches the device) and through rt_scene_hash."""
import copy
import ctypes as C

import numpy as np
import pytest

import mu_lambda_raytracer_b200 as rt
from mu_lambda_raytracer_b200 import abi
import support as S


def mutable_copy(desc):
    """deep copy of an RtSceneDesc into fresh ctypes arrays (returned with the objects that keep them alive)"""
    d = desc.contents if hasattr(desc, "contents") else desc
    keep = {}

    def arr(ctype, src, n):
        a = (ctype * max(n, 1))()
        if n:
            C.memmove(a, src, n * C.sizeof(ctype))
        return a

    out = abi.RtSceneDesc()
    C.memmove(C.byref(out), C.byref(d), C.sizeof(out))
    keep["nodes"] = arr(abi.RtNode, d.nodes, d.n_nodes)
    keep["children"] = arr(C.c_int32, d.children, d.n_children)
    keep["materials"] = arr(abi.RtMaterial, d.materials, d.n_materials)
    keep["textures"] = arr(abi.RtTexture, d.textures, d.n_textures)
    out.nodes, out.children = keep["nodes"], keep["children"]
    out.materials, out.textures = keep["materials"], keep["textures"]
    return out, keep


@pytest.mark.parametrize("name", ["cornell_smoke", "simple_light", "final_scene"])
def test_corrupted_descriptions_are_refused_not_crashed(name):
    base = rt.World(name).build(42)
    rng = np.random.default_rng(sum(map(ord, name)))
    L = S.emul()
    lib = abi.load()
    refused = accepted = 0
    for trial in range(150):
        d, keep = mutable_copy(base.ptr)
        for _ in range(int(rng.integers(1, 4))):
            what = int(rng.integers(0, 9))
            i = int(rng.integers(0, d.n_nodes))
            weird = [-(2 ** 31), -7, -1, 0, 1, 5, 11, 12, 255, 2 ** 20, 2 ** 31 - 1]
            pick = lambda: int(rng.choice(weird))
            if what == 0:
                d.nodes[i].kind = pick()
            elif what == 1:
                d.nodes[i].material = pick()
            elif what == 2:
                d.nodes[i].first_child = pick()
            elif what == 3:
                d.nodes[i].child_count = pick()
            elif what == 4:
                d.nodes[i].axis = pick()
            elif what == 5:
                d.nodes[i].f[int(rng.integers(0, 8))] = float(rng.choice([np.nan, np.inf, -np.inf, 0.0, -1.0, 1e300]))
            elif what == 6 and d.n_children:
                d.children[int(rng.integers(0, d.n_children))] = pick()
            elif what == 7 and d.n_materials:
                m = d.materials[int(rng.integers(0, d.n_materials))]
                m.kind, m.texture = (pick(), m.texture) if rng.random() < 0.5 else (m.kind, pick())
            elif what == 8 and d.n_textures:
                t = d.textures[int(rng.integers(0, d.n_textures))]
                t.kind, t.a, t.b = pick() if rng.random() < 0.3 else t.kind, pick(), pick()
            if rng.random() < 0.1:
                d.root = pick()
        h = L.emul_scene_create(C.byref(d), -1, 1)  # flatten + host BVH build; NULL = refused with a message
        if h:
            accepted += 1
            L.emul_scene_destroy(h)
        else:
            refused += 1
            assert len(L.emul_last_error()) > 0
        digest = (C.c_uint8 * 32)()
        assert lib.rt_scene_hash(C.byref(d), digest) in (abi.RT_OK, abi.RT_ERR_INVALID)
    assert refused > 20 and accepted > 5, (refused, accepted)  # both outcomes were exercised
