"""Host-side mirror of the reference's render interface over the C ABI (no compute happens in Python).

Reference surface being mirrored (names, argument meaning, error behaviour):
  worlds() / World{name, camera, background, build}      src/worlds.rs:14-19, :471-484
  Camera::new(lookfrom, lookat, vup, vfov, aspect, aperture, focus_dist)   src/camera.rs:15-38
  RenderingParams{samples_per_pixel, image_height, image_width}           src/raytrace.rs:50-55
  RecursiveRayTracer{max_depth}                                           src/raytrace.rs:74-76
  SeedableRngator::new(seed)                                              src/rngator.rs:17-25
  Renderer::new_with_rng(camera, world, background, params, tracer, rng).render(logger) -> rows of (r,g,b),
      row j = 0 is the BOTTOM row, logger(j, height) called `height` times   src/raytrace.rs:151-186
"""
import ctypes as C
import os

import numpy as np

from . import abi

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_EARTHMAP = os.path.join(_ROOT, "assets", "earthmap.ppm")


def load_earthmap(path=None):
    """Decoded RGB8 texels of the reference's earthmap.jpg (1024x512), shipped as a binary PPM so that every host
    sees the same bytes (jpeg decoders differ by +-1 LSB; SURVEY hard part 8).  Returns (H, W, 3) uint8."""
    path = path or os.environ.get("RT_EARTHMAP") or DEFAULT_EARTHMAP
    with open(path, "rb") as f:
        data = f.read()
    if data[:2] != b"P6":
        raise ValueError(f"{path}: expected a binary PPM (P6)")
    fields, pos = [], 2
    while len(fields) < 3:
        while data[pos:pos + 1].isspace():
            pos += 1
        if data[pos:pos + 1] == b"#":
            pos = data.index(b"\n", pos)
            continue
        end = pos
        while not data[end:end + 1].isspace():
            end += 1
        fields.append(int(data[pos:end]))
        pos = end
    w, h, maxval = fields
    if maxval != 255:
        raise ValueError("earthmap PPM must be 8-bit")
    return np.frombuffer(data, dtype=np.uint8, offset=pos + 1, count=w * h * 3).reshape(h, w, 3).copy()


class GradientBackground:  # raytrace.rs:12-35
    kind = abi.RT_BG_GRADIENT


class BlackBackground:  # raytrace.rs:37-48
    kind = abi.RT_BG_BLACK


class RenderingParams:  # raytrace.rs:50-55
    def __init__(self, samples_per_pixel, image_height, image_width):
        self.samples_per_pixel = int(samples_per_pixel)
        self.image_height = int(image_height)
        self.image_width = int(image_width)


class RecursiveRayTracer:  # raytrace.rs:74-76
    def __init__(self, max_depth=50):
        self.max_depth = int(max_depth)


class SeedableRngator:  # rngator.rs:17-31 — on the device the seed keys the Philox render streams
    def __init__(self, seed):
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF


class Camera:  # camera.rs:15-38: the seven inputs; the basis is derived inside the library in f64
    def __init__(self, lookfrom, lookat, vup, vfov, aspect_ratio, aperture, focus_dist, time0=0.0, time1=0.0):
        """time0/time1: shutter interval in [0, 1] of the motion-blur EXTENSION (the reference's camera has none: 0, 0)"""
        self.c = abi.RtCamera()
        for i in range(3):
            self.c.lookfrom[i], self.c.lookat[i], self.c.vup[i] = float(lookfrom[i]), float(lookat[i]), float(vup[i])
        self.c.vfov_deg, self.c.aspect_ratio = float(vfov), float(aspect_ratio)
        self.c.aperture, self.c.focus_dist = float(aperture), float(focus_dist)
        self.c.time0, self.c.time1 = float(time0), float(time1)


class SceneDescription:
    """What World::build produces here: the reference's object tree written down as an RtSceneDesc."""

    def __init__(self, ptr, n_draws=0, owned=True):
        self.ptr, self.n_draws, self._owned = ptr, n_draws, owned

    @property
    def desc(self):
        return self.ptr.contents

    def hash(self):
        out = (C.c_uint8 * 32)()
        abi.check(abi.load().rt_scene_hash(self.ptr, out))
        return bytes(out).hex()

    def __del__(self):
        try:
            if getattr(self, "_owned", False) and self.ptr:
                abi.load().rt_scene_desc_free(self.ptr)
                self.ptr = None
        except Exception:  # interpreter shutdown: the module globals may already be gone
            pass


class Scene:
    """A description flattened, BVH-built and resident on one GPU (rt_scene_create)."""

    def __init__(self, description, device=-1):
        self.description = description  # keep alive for sub-tree queries by the caller
        h = C.c_void_p()
        abi.check(abi.load().rt_scene_create(description.ptr, int(device), C.byref(h)))
        self.handle = h

    def info(self):
        a, b, c, d = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
        abi.check(abi.load().rt_scene_info(self.handle, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return {"prims": a.value, "bvh_nodes": b.value, "media": c.value, "device_bytes": d.value}

    def build_info(self):
        a, b, c = C.c_int32(), C.c_float(), C.c_int32()
        abi.check(abi.load().rt_scene_build_info(self.handle, C.byref(a), C.byref(b), C.byref(c)))
        return {"built_on_device": bool(a.value), "build_ms": b.value, "depth": c.value}

    def close(self):
        if self.handle:
            abi.load().rt_scene_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class World:  # worlds.rs:14-19
    def __init__(self, name):
        self._name = name
        self._info = abi.RtWorldInfo()
        abi.check(abi.load().rt_world_info(name.encode(), C.byref(self._info)))

    def name(self):
        return self._name

    def camera(self):
        return {"lookfrom": tuple(self._info.lookfrom), "lookat": tuple(self._info.lookat),
                "field_of_view": self._info.vfov_deg}

    def background(self):
        return GradientBackground() if self._info.background_kind == abi.RT_BG_GRADIENT else BlackBackground()

    def build(self, seed, earthmap=None):
        """World::build with rng = Pcg64::seed_from_u64(seed) (main.rs:185); returns a SceneDescription."""
        ptr, draws = C.POINTER(abi.RtSceneDesc)(), C.c_uint64()
        earth, ew, eh = None, 0, 0
        if self._info.needs_earthmap:
            img = load_earthmap() if earthmap is None else np.ascontiguousarray(earthmap, dtype=np.uint8)
            eh, ew = img.shape[:2]
            earth = img.ctypes.data_as(C.c_void_p)
            self._keep = img
        abi.check(abi.load().rt_world_build(self._name.encode(), int(seed) & 0xFFFFFFFFFFFFFFFF, earth, ew, eh,
                                            C.byref(ptr), C.byref(draws)))
        return SceneDescription(ptr, draws.value)


def worlds():  # worlds.rs:471-484, same order
    lib = abi.load()
    return [World(lib.rt_world_name(i).decode()) for i in range(lib.rt_world_count())]


class Renderer:  # raytrace.rs:137-198
    def __init__(self, camera, world, background, parameters, tracer, rng):
        if not isinstance(world, Scene):
            raise TypeError("world must be a Scene (a description resident on the GPU)")
        self.camera, self.world, self.background = camera, world, background
        self.parameters, self.tracer, self.rng = parameters, tracer, rng
        self.stats = None
        self.pipeline = abi.RT_PIPELINE_AUTO
        self.bvh_layout = 0  # 0 = auto, 2 = binary 32-byte-node BVH, 4 = 4-wide BVH (RtParams.bvh_layout)

    @classmethod
    def new_with_rng(cls, camera, world, background, parameters, tracer, rng):
        return cls(camera, world, background, parameters, tracer, rng)

    def _params(self):
        p = abi.RtParams()
        p.width, p.height = self.parameters.image_width, self.parameters.image_height
        p.samples_per_pixel, p.max_depth = self.parameters.samples_per_pixel, self.tracer.max_depth
        p.seed, p.sample_begin, p.sample_count = self.rng.seed, 0, 0
        p.pipeline, p.device, p.bvh_layout = self.pipeline, -1, self.bvh_layout
        return p

    def render_arrays(self, logger=None, want_accum=True):
        """One rt_render call with HOST buffers.  Returns (rgb int32 [H,W,3], accum float32 [H,W,3] or None);
        row 0 is the bottom row.  `logger(j, H)` is called on this thread once per row index, as Renderer::render does
        (raytrace.rs:182), paced by the device's progress."""
        p = self._params()
        h, w = p.height, p.width
        rgb = np.empty((h, w, 3), dtype=np.int32)
        accum = np.empty((h, w, 3), dtype=np.float32) if want_accum else None
        cb = abi.RtProgressFn(lambda row, total, user: logger(row, total)) if logger else abi.RtProgressFn()
        stats = abi.RtStats()
        abi.check(abi.load().rt_render(self.world.handle, C.byref(self.camera.c), C.byref(p),
                                       accum.ctypes.data_as(C.c_void_p) if want_accum else None,
                                       rgb.ctypes.data_as(C.c_void_p), cb, None, C.byref(stats)))
        self.stats = {"paths": stats.paths, "rays": stats.rays, "device_ms": stats.device_ms,
                      "kernel_launches": stats.kernel_launches, "pipeline": stats.pipeline_used,
                      "bvh_layout": stats.bvh_layout_used}
        return rgb, accum

    def render(self, logger=None):
        """Renderer::render: H rows of W (r, g, b) tuples, row j = 0 at the BOTTOM; logger(j, H) called H times."""
        rgb, _ = self.render_arrays(logger, want_accum=False)
        return [[tuple(int(c) for c in px) for px in row] for row in rgb]


def to_ppm(rgb):
    """main.rs:144,175-179: 'P3\\nW H\\n255' then one 'r g b' line per pixel, rows in reverse j (top row first)."""
    h, w = rgb.shape[:2]
    lines = ["P3", f"{w} {h}", "255"]
    for j in range(h - 1, -1, -1):
        lines.extend(f"{int(r)} {int(g)} {int(b)}" for r, g, b in rgb[j])
    return "\n".join(lines) + "\n"
