"""Shared test plumbing: the oracle (oracle/liboracle.so), the host emulation of the device header
(tests/emul/libemul.so) and small helpers.  Both libraries are TEST infrastructure; the product never loads them."""
import ctypes as C
import os
import subprocess

import numpy as np

import mu_lambda_raytracer_b200 as rt
from mu_lambda_raytracer_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


class OrcHit(C.Structure):
    _fields_ = [("hit", C.c_int32), ("front_face", C.c_int32), ("material", C.c_int32), ("node", C.c_int32),
                ("t", C.c_double), ("u", C.c_double), ("v", C.c_double), ("p", C.c_double * 3),
                ("normal", C.c_double * 3)]


_oracle = None
_emul = None


def oracle():
    global _oracle
    if _oracle is None:
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        srcs = [os.path.join(ROOT, "oracle", f) for f in ("capi.cpp", "pcg64.hpp", "scene.hpp", "worlds.hpp", "render.hpp")]
        if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
        L = C.CDLL(path)
        L.orc_world_build.restype = C.c_void_p
        L.orc_world_build.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p, C.c_int, C.c_int]
        L.orc_world_from_desc.restype = C.c_void_p
        L.orc_world_from_desc.argtypes = [C.POINTER(abi.RtSceneDesc), C.c_uint64]
        L.orc_world_free.argtypes = [C.c_void_p]
        L.orc_world_desc.restype = C.POINTER(abi.RtSceneDesc)
        L.orc_world_desc.argtypes = [C.c_void_p]
        L.orc_world_info.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                     C.POINTER(C.c_int32), C.POINTER(C.c_uint64), C.POINTER(C.c_int32)]
        L.orc_world_bvh_axes.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
        L.orc_render.restype = C.c_double
        L.orc_render.argtypes = [C.c_void_p, C.POINTER(abi.RtCamera), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                 C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_hit_batch.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_uint64, C.c_int32, C.c_void_p]
        L.orc_hit_batch_at.argtypes = [C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_int64, C.c_uint64, C.c_int32, C.c_void_p]
        L.orc_medium_interval_batch.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_pcg_seed_stream.argtypes = [C.c_uint64, C.c_int32, C.c_void_p]
        L.orc_pcg_new_stream.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int32, C.c_void_p]
        L.orc_pcg_from_seed_first.restype = C.c_uint64
        L.orc_pcg_from_seed_first.argtypes = [C.c_void_p]
        L.orc_pcg_f64_stream.argtypes = [C.c_uint64, C.c_double, C.c_double, C.c_int32, C.c_void_p]
        L.orc_pcg_usize_stream.restype = C.c_uint64
        L.orc_pcg_usize_stream.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int32, C.c_void_p]
        L.orc_sphere_uv.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_aabb_hit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double]
        L.orc_aabb_corners.argtypes = [C.c_void_p] * 4
        L.orc_to_rgb.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.orc_camera_ray.argtypes = [C.POINTER(abi.RtCamera), C.c_double, C.c_double, C.c_uint64, C.c_void_p, C.c_void_p]
        L.orc_camera_basis.argtypes = [C.POINTER(abi.RtCamera), C.c_void_p]
        L.orc_texture_value.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_scatter_batch.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_uint64, C.c_void_p]
        L.orc_hardware_threads.restype = C.c_int32
        _oracle = L
    return _oracle


def emul():
    """Device header compiled for the host (tests/emul/emul.cpp); CPU-only debugging aid for the device math."""
    global _emul
    if _emul is None:
        path = os.path.join(ROOT, "tests", "emul", "libemul.so")
        csrc = os.path.join(ROOT, "mu-lambda-raytracer_b200", "csrc")
        srcs = [os.path.join(ROOT, "tests", "emul", "emul.cpp")] + [os.path.join(csrc, f) for f in
                                                                     ("flatten.cpp", "flatten.h", "rt_device.cuh", "rt_types.h")]
        if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-pthread", "-w", "-shared", "-o", path,
                                   srcs[0], srcs[1]])
        L = C.CDLL(path)
        L.emul_last_error.restype = C.c_char_p
        L.emul_scene_create.restype = C.c_void_p
        L.emul_scene_create.argtypes = [C.POINTER(abi.RtSceneDesc), C.c_int32, C.c_int32]
        L.emul_scene_destroy.argtypes = [C.c_void_p]
        L.emul_scene_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int32)] * 4
        L.emul_prim_nodes.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        L.emul_intersect_batch.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]
        L.emul_intersect_batch_at.argtypes = [C.c_void_p, C.c_int32, C.c_float, C.c_void_p, C.c_int64, C.c_void_p]
        L.emul_texture_value_batch.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]
        L.emul_scatter_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.emul_generate_rays.argtypes = [C.POINTER(abi.RtCamera), C.POINTER(abi.RtParams), C.c_void_p, C.c_void_p,
                                         C.c_int64, C.c_void_p, C.c_void_p]
        L.emul_philox.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.emul_unit_ball.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.emul_render.restype = C.c_uint64
        L.emul_render.argtypes = [C.c_void_p, C.POINTER(abi.RtCamera), C.POINTER(abi.RtParams), C.c_int32, C.c_void_p]
        L.emul_clear_media.restype = C.c_uint32
        L.emul_clear_media.argtypes = [C.c_void_p]
        L.emul_scene_features.restype = C.c_uint32
        L.emul_scene_features.argtypes = [C.c_void_p]
        L.emul_render_slots.argtypes = [C.c_void_p, C.POINTER(abi.RtCamera), C.POINTER(abi.RtParams), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        _emul = L
    return _emul


_earth = None


def earthmap():
    global _earth
    if _earth is None:
        _earth = rt.load_earthmap()
    return _earth


class OracleWorld:
    def __init__(self, name=None, seed=42, desc=None, bvh_seed=1):
        L = oracle()
        if desc is not None:
            self._keep = desc
            self.h = L.orc_world_from_desc(desc, bvh_seed)
        else:
            e = earthmap()
            self.h = L.orc_world_build(name.encode(), seed, e.ctypes.data, e.shape[1], e.shape[0])
        if not self.h:
            raise RuntimeError("oracle could not build the world")
        lf, la = (C.c_double * 3)(), (C.c_double * 3)()
        fov, bg, dr, nb = C.c_double(), C.c_int32(), C.c_uint64(), C.c_int32()
        L.orc_world_info(self.h, lf, la, C.byref(fov), C.byref(bg), C.byref(dr), C.byref(nb))
        self.lookfrom, self.lookat, self.vfov = tuple(lf), tuple(la), fov.value
        self.background, self.draws, self.n_bvh = bg.value, dr.value, nb.value

    @property
    def desc(self):
        return oracle().orc_world_desc(self.h)

    def bvh_axes(self, which):
        L = oracle()
        n = L.orc_world_bvh_axes(self.h, which, None, 0)
        out = np.zeros(n, dtype=np.int32)
        L.orc_world_bvh_axes(self.h, which, out.ctypes.data, n)
        return out

    def hit(self, rays, node=-1, rng_seed=7, skip_media=True, time=0.0):
        rays = np.ascontiguousarray(rays, dtype=np.float64)
        out = (OrcHit * len(rays))()
        rc = oracle().orc_hit_batch_at(self.h, node, float(time), rays.ctypes.data, len(rays), rng_seed, 1 if skip_media else 0, out)
        assert rc == 0
        return np.ctypeslib.as_array(out)

    def medium_interval(self, rays, node):
        rays = np.ascontiguousarray(rays, dtype=np.float64)
        hit = np.zeros(len(rays), dtype=np.int32)
        t = np.zeros((len(rays), 2), dtype=np.float64)
        rc = oracle().orc_medium_interval_batch(self.h, node, rays.ctypes.data, len(rays), hit.ctypes.data, t.ctypes.data)
        assert rc == 0
        return hit, t

    def render(self, cam, width, height, spp, max_depth=50, render_seed=42, rows=None, threads=0, want_rgb=True):
        accum = np.zeros((height, width, 3), dtype=np.float64)
        rgb = np.zeros((height, width, 3), dtype=np.int32)
        counters = np.zeros(16, dtype=np.uint64)
        r0, r1 = rows if rows else (0, -1)
        secs = oracle().orc_render(self.h, C.byref(cam), width, height, spp, max_depth, render_seed, r0, r1, threads,
                                   accum.ctypes.data, rgb.ctypes.data, counters.ctypes.data)
        return accum, rgb, counters, secs

    def close(self):
        if self.h:
            oracle().orc_world_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


COUNTER_NAMES = ["paths", "rays", "aabb", "sphere", "rect", "xform", "medium", "lambertian", "metal", "dielectric",
                 "light", "isotropic", "perlin", "image", "background", "depth_exhausted"]


def make_camera(lookfrom, lookat, vfov, aspect, aperture=0.0, focus_dist=None, vup=(0, 1, 0), time0=0.0, time1=0.0):
    if focus_dist is None:  # main.rs:109-112
        focus_dist = float(np.linalg.norm(np.asarray(lookat, float) - np.asarray(lookfrom, float)))
    return rt.Camera(lookfrom, lookat, vup, vfov, aspect, aperture, focus_dist, time0, time1)


class EmulScene:
    def __init__(self, desc_ptr, root=-1, build_bvh=True):
        L = emul()
        self._keep = desc_ptr
        self.h = L.emul_scene_create(desc_ptr, root, 1 if build_bvh else 0)
        if not self.h:
            raise RuntimeError("emul: " + L.emul_last_error().decode())
        a, b, c, d = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        L.emul_scene_info(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        self.n_prims, self.n_nodes, self.n_media, self.depth = a.value, b.value, c.value, d.value
        self.prim_nodes = np.zeros(max(self.n_prims, 1), dtype=np.int32)
        L.emul_prim_nodes(self.h, self.prim_nodes.ctypes.data, self.n_prims)

    def intersect(self, rays, mode=0, time=0.0):
        rays = np.ascontiguousarray(rays, dtype=np.float32)
        out = (abi.RtHit * len(rays))()
        emul().emul_intersect_batch_at(self.h, mode, float(time), rays.ctypes.data, len(rays), out)
        a = np.ctypeslib.as_array(out).copy()
        ok = a["prim"] >= 0
        a["prim"][ok] = self.prim_nodes[a["prim"][ok]]
        return a

    def render(self, cam, width, height, spp, max_depth=50, seed=42, threads=0, sample_begin=0, fixed=False):
        """radiance sums per pixel: float32 (as rt_render's accum_rgb), or with fixed=True the library's native 64-bit
        fixed-point sums (2^-32 units) as uint64"""
        p = abi.RtParams()
        p.width, p.height, p.samples_per_pixel, p.max_depth = width, height, spp, max_depth
        p.seed, p.sample_begin, p.sample_count = seed, sample_begin, spp
        accum = np.zeros((height, width, 3), dtype=np.uint64)
        rays = emul().emul_render(self.h, C.byref(cam.c), C.byref(p), threads, accum.ctypes.data)
        if fixed:
            return accum, rays
        return (accum.astype(np.float64) / abi.RT_ACCUM_FIXED_ONE).astype(np.float32), rays

    @property
    def clear_media(self):
        return int(emul().emul_clear_media(self.h))

    @property
    def features(self):
        return int(emul().emul_scene_features(self.h))

    def render_slots(self, cam, width, height, spp, chain, max_depth=50, seed=42, threads=0):
        """the wavefront slot functions driven path by path (fixed-point sums, rays, chain steps); chain=True advances
        paths inside a clear medium with wf_chain_step as the persistent kernel's chain phase does"""
        p = abi.RtParams()
        p.width, p.height, p.samples_per_pixel, p.max_depth = width, height, spp, max_depth
        p.seed, p.sample_begin, p.sample_count = seed, 0, spp
        accum = np.zeros((height, width, 3), dtype=np.uint64)
        counters = np.zeros(2, dtype=np.uint64)
        emul().emul_render_slots(self.h, C.byref(cam.c), C.byref(p), threads, 1 if chain else 0, accum.ctypes.data, counters.ctypes.data)
        return accum, int(counters[0]), int(counters[1])

    def close(self):
        if self.h:
            emul().emul_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def random_rays(n, rng, lo, hi, tmin=0.001, tmax=np.inf, target=None, spread=None):
    """n rays with origins uniform in the box [lo, hi] and NON-unit directions; with `target`, directions aim at
    a point jittered by `spread` around it (so that small objects get hit often)."""
    o = rng.uniform(lo, hi, size=(n, 3))
    if target is None:
        d = rng.normal(size=(n, 3))
    else:
        aim = np.asarray(target, float) + rng.uniform(-1, 1, size=(n, 3)) * np.asarray(spread, float)
        d = aim - o
    d *= rng.uniform(0.05, 3.0, size=(n, 1)) / np.linalg.norm(d, axis=1, keepdims=True)
    r = np.zeros((n, 8))
    r[:, 0:3], r[:, 3:6], r[:, 6], r[:, 7] = o, d, tmin, tmax
    # the device sees f32 rays: make the f64 oracle see exactly the same numbers
    return r.astype(np.float32).astype(np.float64)


class DescBuilder:
    """Builds ad-hoc RtSceneDesc's for the per-primitive parity tests (the same struct the product builder emits)."""

    def __init__(self):
        self.nodes, self.children, self.materials, self.textures, self.perlins, self.images = [], [], [], [], [], []
        self._keep = []

    def solid(self, r, g, b):
        t = abi.RtTexture()
        t.kind, t.a, t.b = abi.RT_TEX_SOLID, -1, -1
        t.color[0], t.color[1], t.color[2] = r, g, b
        self.textures.append(t)
        return len(self.textures) - 1

    def checker(self, odd, even):
        t = abi.RtTexture()
        t.kind, t.a, t.b = abi.RT_TEX_CHECKER, odd, even
        self.textures.append(t)
        return len(self.textures) - 1

    def noise(self, scale, perlin_from=None, seed=3):
        """perlin tables: copied from another description, or random unit vectors + permutations"""
        p = abi.RtPerlin()
        if perlin_from is not None:
            C.memmove(C.byref(p), C.byref(perlin_from), C.sizeof(p))
        else:
            rng = np.random.default_rng(seed)
            v = rng.normal(size=(1024, 3))
            v /= np.linalg.norm(v, axis=1, keepdims=True)
            for i in range(1024):
                for c in range(3):
                    p.ranvec[i][c] = v[i, c]
            for perm in (p.perm_x, p.perm_y, p.perm_z):
                for i, x in enumerate(rng.permutation(1024)):
                    perm[i] = int(x)
        self.perlins.append(p)
        t = abi.RtTexture()
        t.kind, t.a, t.b, t.scale = abi.RT_TEX_NOISE, len(self.perlins) - 1, -1, scale
        self.textures.append(t)
        return len(self.textures) - 1

    def image(self, rgb):
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        self._keep.append(rgb)
        im = abi.RtImage()
        im.width, im.height = rgb.shape[1], rgb.shape[0]
        im.rgb = rgb.ctypes.data_as(C.POINTER(C.c_uint8))
        self.images.append(im)
        t = abi.RtTexture()
        t.kind, t.a, t.b = abi.RT_TEX_IMAGE, len(self.images) - 1, -1
        self.textures.append(t)
        return len(self.textures) - 1

    def material(self, kind, tex=-1, albedo=(0, 0, 0), fuzz=0.0, ior=0.0):
        m = abi.RtMaterial()
        m.kind, m.texture, m.fuzz, m.ior = kind, tex, fuzz, ior
        for i in range(3):
            m.albedo[i] = albedo[i]
        self.materials.append(m)
        return len(self.materials) - 1

    def lambertian(self, tex):
        return self.material(abi.RT_MAT_LAMBERTIAN, tex)

    def _node(self, kind, mat, f, child=-1, axis=0, count=0):
        n = abi.RtNode()
        n.kind, n.material, n.first_child, n.child_count, n.axis = kind, mat, child, count, axis
        for i, v in enumerate(f):
            n.f[i] = v
        self.nodes.append(n)
        return len(self.nodes) - 1

    def sphere(self, c, r, mat):
        return self._node(abi.RT_NODE_SPHERE, mat, [c[0], c[1], c[2], r])

    def moving_sphere(self, c0, c1, r, mat):  # extension: centre(time) = c0 + time * (c1 - c0)
        return self._node(abi.RT_NODE_MOVING_SPHERE, mat, [c0[0], c0[1], c0[2], c1[0], c1[1], c1[2], r])

    def rect(self, kind, a0, a1, b0, b1, k, mat):
        return self._node(kind, mat, [a0, a1, b0, b1, k])

    def block(self, p0, p1, mat):
        return self._node(abi.RT_NODE_BLOCK, mat, list(p0) + list(p1))

    def translate(self, off, child):
        return self._node(abi.RT_NODE_TRANSLATE, -1, list(off), child)

    def rotate(self, axis, degrees, child):
        return self._node(abi.RT_NODE_ROTATE, -1, [degrees], child, axis)

    def medium(self, boundary, density, color):
        iso = self.material(abi.RT_MAT_ISOTROPIC, self.solid(*color))
        return self._node(abi.RT_NODE_MEDIUM, iso, [density], boundary)

    def group(self, kind, items):
        first = len(self.children)
        self.children.extend(items)
        return self._node(kind, -1, [], first, 0, len(items))

    def finish(self, root, background=abi.RT_BG_BLACK):
        d = abi.RtSceneDesc()
        d.root, d.background_kind = root, background
        if background == abi.RT_BG_GRADIENT:
            for i, (t, b) in enumerate(zip((0.5, 0.7, 1.0), (1.0, 1.0, 1.0))):
                d.background_top[i], d.background_bottom[i] = t, b

        def arr(ctype, items):
            a = (ctype * max(len(items), 1))(*items)
            self._keep.append(a)
            return a

        d.n_nodes, d.nodes = len(self.nodes), arr(abi.RtNode, self.nodes)
        d.n_children, d.children = len(self.children), arr(C.c_int32, self.children)
        d.n_materials, d.materials = len(self.materials), arr(abi.RtMaterial, self.materials)
        d.n_textures, d.textures = len(self.textures), arr(abi.RtTexture, self.textures)
        d.n_perlins, d.perlins = len(self.perlins), arr(abi.RtPerlin, self.perlins)
        d.n_images, d.images = len(self.images), arr(abi.RtImage, self.images)
        self.desc = d
        return C.pointer(d)
