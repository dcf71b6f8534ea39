#!/usr/bin/env python
"""Algorithmic work of the REFERENCE's traversal per camera path, counted by the oracle's instrumentation
on the reference's own BVH topology: the frozen figures behind bench.py's roofline and BASELINE.md §4.
usage: python tests/golden/algo_work.py [spp]   (full-resolution frames at reduced spp; cost is linear in spp)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import support as S  # noqa: E402

CONFIGS = [("C1", "random", 400, 266, 1.5, {}), ("C2", "random", 1200, 800, 1.5, dict(aperture=0.1, focus_dist=10.0)),
           ("C3", "cornell_smoke", 600, 600, 1.0, {}), ("C4", "final_scene", 800, 800, 1.0, {})]
NAMES = ["paths", "rays", "aabb", "sphere", "rect", "xform", "medium", "lambertian", "metal", "dielectric", "light", "isotropic",
         "perlin", "image", "background", "depth_exhausted"]


def main(spp):
    out = {}
    for tag, world, w, h, aspect, kw in CONFIGS:
        ow = S.OracleWorld(world, 42)
        cam = S.make_camera(ow.lookfrom, ow.lookat, ow.vfov, aspect, **kw)
        _, _, c, secs = ow.render(cam.c, w, h, spp, render_seed=42)
        c = dict(zip(NAMES, (int(x) for x in c)))
        p = c["paths"]
        per = {k: c[k] / p for k in NAMES[1:]}
        geometry = 27 * per["aabb"] + 45 * per["sphere"] + 15 * per["rect"] + 12 * per["xform"] + 40 * per["medium"]
        shade = (40 * per["lambertian"] + 60 * per["metal"] + 80 * per["dielectric"] + 25 * per["isotropic"] + 1400 * per["perlin"]
                 + 55 * per["image"] + 20 * per["background"])
        per["flops_per_path"] = geometry + shade
        per["bytes_per_path"] = 32 * (per["aabb"] + per["sphere"] + per["rect"]) + 48 * per["medium"]
        per["cpu_mpaths_per_s"] = p / secs / 1e6
        per["cpu_threads"] = int(S.oracle().orc_hardware_threads())
        per["sample"] = f"{w}x{h} at {spp} spp"
        out[tag] = per
        print(f"{tag} {world:14s} rays/path {per['rays']:.3f}  aabb/ray {per['aabb'] / per['rays']:.1f}  prim/ray "
              f"{(per['sphere'] + per['rect']) / per['rays']:.1f}  flops/path {per['flops_per_path']:.0f}  bytes/path {per['bytes_per_path']:.0f}  "
              f"cpu {per['cpu_mpaths_per_s']:.2f} Mpaths/s on {per['cpu_threads']} threads", flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 8)
