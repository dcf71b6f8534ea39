#!/usr/bin/env python
"""Generates the committed golden fixtures under tests/golden/.  Run in the build container (it reads the
reference's published renders under /root/reference, which do not exist on the GPU box; the tests only read the
JSON this script writes).

1. rng_golden.json — an INDEPENDENT pure-Python big-integer restatement of the un-vendored RNG crates
   (rand_core 0.6.2 seed_from_u64, rand_pcg 0.3.0 Lcg128Xsl64, rand 0.8.3 gen_range) and the values the three
   config worlds draw from it for seed 42.  Checked against upstream's published known answers before writing.
2. published_render_probe.json — colours of the reference's own README renders (sample.jpg, final_scene.jpg) at
   the projected centres of scene objects: the only evidence about the reference's output that exists
   independently of any restatement; for final_scene.jpg the image box that the 1000 foam spheres
   (the last objects the world RNG places) project into, from this file's own restatement of the recipe.
3. published_final_scene.png, published_sample_blur.png — the decoded pixels of the reference's README renders
   (final_scene.jpg, sample_blur.jpg; README.md:20-38 of the reference), stored losslessly: what the image-level
   tests compare this repo's renders of the same command lines with.
"""
import json
import math
import os
import struct

HERE = os.path.dirname(os.path.abspath(__file__))
M64, M128 = (1 << 64) - 1, (1 << 128) - 1
PCG_MULT = 0x2360ED051FC65DA44385DF649FCCF645


class Pcg64:
    def __init__(self, state, incr):
        self.incr = incr & M128
        self.state = (state + self.incr) & M128
        self.calls = 0
        self._step()

    def _step(self):
        self.state = (self.state * PCG_MULT + self.incr) & M128

    @classmethod
    def new(cls, state, stream):
        return cls(state, ((stream << 1) | 1) & M128)

    @classmethod
    def from_seed(cls, seed32):
        s = struct.unpack("<4Q", bytes(seed32))
        return cls(s[0] | (s[1] << 64), (s[2] | (s[3] << 64)) | 1)

    @classmethod
    def seed_from_u64(cls, x):
        out = b""
        for _ in range(8):
            x = (x * 6364136223846793005 + 11634580027462260723) & M64
            xs = (((x >> 18) ^ x) >> 27) & 0xFFFFFFFF
            rot = x >> 59
            out += struct.pack("<I", ((xs >> rot) | (xs << ((32 - rot) & 31))) & 0xFFFFFFFF)
        return cls.from_seed(out)

    def next_u64(self):
        self.calls += 1
        self._step()
        rot = self.state >> 122
        x = ((self.state >> 64) ^ self.state) & M64
        return ((x >> rot) | (x << ((64 - rot) & 63))) & M64

    def f64(self, lo, hi):
        scale = hi - lo
        while True:
            bits = (self.next_u64() >> 12) | 0x3FF0000000000000
            v = struct.unpack("<d", struct.pack("<Q", bits))[0] - 1.0
            r = v * scale + lo
            if r < hi:
                return r
            scale = struct.unpack("<d", struct.pack("<Q", struct.unpack("<Q", struct.pack("<d", scale))[0] - 1))[0]

    def usize(self, lo, hi):
        rng = hi - lo
        zone = ((rng << (64 - rng.bit_length())) - 1) & M64
        while True:
            m = self.next_u64() * rng
            if (m & M64) <= zone:
                return lo + (m >> 64)


def check_upstream_known_answers():
    g = Pcg64.new(42, 54)
    assert [g.next_u64() for _ in range(6)] == [0x86b1da1d72062b68, 0x1304aa46c9853d39, 0xa3670e9e0dd50358,
                                               0xf9090e529a7dae00, 0xc85b9fd837996f2c, 0x606121f8e3919196]
    assert Pcg64.from_seed(range(1, 33)).next_u64() == 8740028313290271629


def bvh_axes(rng, n, out):
    if n < 2:
        return
    out.append(rng.usize(0, 3))
    bvh_axes(rng, n // 2, out)
    bvh_axes(rng, n - n // 2, out)


def random_world(seed):
    r = Pcg64.seed_from_u64(seed)
    spheres = []
    for a in range(-11, 11):
        for b in range(-11, 11):
            choose = r.f64(0.0, 1.0)
            cx = a + 0.9 * r.f64(0.0, 1.0)
            cz = b + 0.9 * r.f64(0.0, 1.0)
            if math.sqrt((cx - 4.0) ** 2 + 0.0 + cz ** 2) > 0.9:
                if choose < 0.8:
                    p = [r.f64(0.0, 1.0) for _ in range(3)]
                    q = [r.f64(0.0, 1.0) for _ in range(3)]
                    spheres.append(("lambertian", (cx, 0.2, cz), [x * y for x, y in zip(p, q)]))
                elif choose < 0.95:
                    alb = [r.f64(0.5, 1.0) for _ in range(3)]
                    spheres.append(("metal", (cx, 0.2, cz), alb + [r.f64(0.0, 0.5)]))
                else:
                    spheres.append(("dielectric", (cx, 0.2, cz), []))
    before = r.calls
    axes = []
    bvh_axes(r, len(spheres) + 4, axes)
    return spheres, before, r.calls, axes


def final_scene(seed):
    r = Pcg64.seed_from_u64(seed)
    out = {}
    heights = [r.f64(1.0, 70.0) for _ in range(400)]
    c0 = r.calls
    axes_ground = []
    bvh_axes(r, 400, axes_ground)
    c1 = r.calls
    ranvec = []
    for _ in range(1024):
        v = [r.f64(-1.0, 1.0) for _ in range(3)]
        l = math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
        ranvec.append([x / l for x in v])
    perms = []
    for _ in range(3):
        p = list(range(1024))
        for i in range(1023, 0, -1):
            j = r.usize(0, i)
            p[i], p[j] = p[j], p[i]
        perms.append(p)
    c2 = r.calls
    foam = [[r.f64(0.0, 165.0) for _ in range(3)] for _ in range(1000)]
    c3 = r.calls
    axes_foam = []
    bvh_axes(r, 1000, axes_foam)
    out.update(heights_first4=heights[:4], height_399=heights[399], ground_axes_first10=axes_ground[:10],
               ranvec0=ranvec[0], perm_x_first8=perms[0][:8], perm_y_first8=perms[1][:8], perm_z_first8=perms[2][:8],
               foam0=foam[0], foam999=foam[999], foam_axes_first10=axes_foam[:10],
               calls={"heights": c0, "ground_bvh": c1 - c0, "perlin": c2 - c1, "foam": c3 - c2, "foam_bvh": r.calls - c3,
                      "total": r.calls})
    return out, foam


def camera_project(lookfrom, lookat, vfov, aspect, W, H, p):
    """pixel (i, j_from_top) where the reference's Camera (camera.rs:15-38) sees point p through a pinhole"""
    sub = lambda a, b: [x - y for x, y in zip(a, b)]
    dot = lambda a, b: sum(x * y for x, y in zip(a, b))
    cross = lambda a, b: [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]
    unit = lambda a: [x / math.sqrt(dot(a, a)) for x in a]
    w = unit(sub(lookfrom, lookat))
    u = unit(cross([0, 1, 0], w))
    v = cross(w, u)
    d = sub(p, lookfrom)
    depth = -dot(d, w)
    if depth <= 0:
        return None
    h = math.tan(math.radians(vfov) / 2)
    s = dot(d, u) / depth / (2 * h * aspect) + 0.5
    t = dot(d, v) / depth / (2 * h) + 0.5
    return s * (W - 1), (1 - t) * (H - 1), depth


def foam_image_box(foam):
    """image box (x0, x1, y0, y1; y from the top) of final_scene's foam: 1000 spheres of radius 10, rotate_y(15) then
    translate(-100, 270, 395) (worlds.rs:455-465), seen by the (478,278,-600) -> (278,278,0), vfov 40 camera"""
    th = math.radians(15.0)
    c, s = math.cos(th), math.sin(th)
    xs, ys = [], []
    for x, y, z in foam:
        p = [c * x + s * z - 100.0, y + 270.0, -s * x + c * z + 395.0]
        px, py, depth = camera_project([478, 278, -600], [278, 278, 0], 40.0, 1.0, 800, 800, p)
        r = 10.0 / depth / (2 * math.tan(math.radians(20.0))) * 799
        xs += [px - r, px + r]
        ys += [py - r, py + r]
    return [int(min(xs)), int(max(xs)) + 1, int(min(ys)), int(max(ys)) + 1]


def copy_published_renders():
    from PIL import Image
    for src, dst in (("final_scene.jpg", "published_final_scene.png"), ("sample_blur.jpg", "published_sample_blur.png")):
        Image.open(os.path.join("/root/reference", src)).convert("RGB").save(os.path.join(HERE, dst), optimize=True)


def probe_published_renders(spheres):
    from PIL import Image
    import numpy as np
    ref = "/root/reference"
    out = {}
    img = np.asarray(Image.open(os.path.join(ref, "sample.jpg")).convert("RGB")).astype(float) / 255.0
    H, W = img.shape[:2]
    lookfrom, lookat = [13, 2, 3], [0, 0, 0]
    cand = []
    for idx, (kind, c, params) in enumerate(spheres):
        if kind != "lambertian":
            continue
        pr = camera_project(lookfrom, lookat, 20.0, 1.5, W, H, c)
        if not pr:
            continue
        x, y, depth = pr
        rad_px = 0.2 / depth / (2 * math.tan(math.radians(10.0))) * (H - 1)
        if rad_px < 4 or not (8 <= x < W - 8 and 8 <= y < H - 8):
            continue
        # unoccluded: no other sphere centre projects within 1.6 radii and is closer
        occluded = False
        for k2, (_, c2, _) in enumerate(spheres):
            if k2 == idx:
                continue
            p2 = camera_project(lookfrom, lookat, 20.0, 1.5, W, H, c2)
            if p2 and p2[2] < depth and math.hypot(p2[0] - x, p2[1] - y) < 1.6 * rad_px:
                occluded = True
                break
        for bc in ([0, 1, 0], [-4, 1, 0], [4, 1, 0]):
            p2 = camera_project(lookfrom, lookat, 20.0, 1.5, W, H, bc)
            r2 = 1.0 / p2[2] / (2 * math.tan(math.radians(10.0))) * (H - 1)
            if p2[2] < depth and math.hypot(p2[0] - x, p2[1] - y) < r2 + 1.0 * rad_px:
                occluded = True
        if occluded:
            continue
        xi, yi = int(round(x)), int(round(y))
        patch = img[yi - 1:yi + 2, xi - 1:xi + 2].reshape(-1, 3).mean(axis=0)
        cand.append({"sphere": idx, "center": list(c), "albedo": params, "pixel": [x, y], "image_rgb_gamma": patch.tolist()})
    out["sample_jpg"] = {"width": W, "height": H, "probes": cand}
    return out


def main():
    check_upstream_known_answers()
    g = Pcg64.seed_from_u64(42)
    raw = [g.next_u64() for _ in range(4)]
    g = Pcg64.seed_from_u64(42)
    unit = [g.f64(0.0, 1.0) for _ in range(3)]
    spheres, before, total, axes = random_world(42)
    kinds = [s[0] for s in spheres]
    fs, foam = final_scene(42)
    rows = {}
    for j in (0, 1, 799):
        g = Pcg64.seed_from_u64(42 + j)
        rows[str(j)] = [g.f64(0.0, 1.0), g.f64(0.0, 1.0)]
    golden = {
        "upstream_kat": {"new_42_54": ["0x%016x" % x for x in (0x86b1da1d72062b68, 0x1304aa46c9853d39, 0xa3670e9e0dd50358,
                                                              0xf9090e529a7dae00, 0xc85b9fd837996f2c, 0x606121f8e3919196)],
                         "from_seed_1_32_first": 8740028313290271629},
        "seed42_raw": ["0x%016x" % x for x in raw], "seed42_unit": unit,
        "random": {"n_small": len(spheres), "lambertian": kinds.count("lambertian"), "metal": kinds.count("metal"),
                   "dielectric": kinds.count("dielectric"), "first_center": list(spheres[0][1]), "first_albedo": spheres[0][2],
                   "last_center": list(spheres[-1][1]), "calls_before_bvh": before, "calls_total": total, "bvh_axes_first10": axes[:10],
                   "n_inner": len(axes)},
        "final_scene": fs, "row_jitter": rows,
    }
    with open(os.path.join(HERE, "rng_golden.json"), "w") as f:
        json.dump(golden, f, indent=1)
    probes = probe_published_renders(spheres)
    probes["final_scene_jpg"] = {"width": 800, "height": 800, "foam_box_x0_x1_y0_y1": foam_image_box(foam)}
    copy_published_renders()
    with open(os.path.join(HERE, "published_render_probe.json"), "w") as f:
        json.dump(probes, f, indent=1)
    print("wrote rng_golden.json and published_render_probe.json;", len(probes["sample_jpg"]["probes"]), "probes")


if __name__ == "__main__":
    main()
