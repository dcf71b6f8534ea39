#!/usr/bin/env python
"""C1 (random, 400x266, 50 spp: 5.32 M paths, ~3 ms of work) rendered repeatedly through rt_render: device time of the
first (cold: the driver loads the kernel) and of the following calls."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mu_lambda_raytracer_b200 as rt
from mu_lambda_raytracer_b200 import abi

world = rt.World("random")
scene = rt.Scene(world.build(42))
info = world.camera()
focus = float(np.linalg.norm(np.asarray(info["lookat"]) - np.asarray(info["lookfrom"])))
cam = rt.Camera(info["lookfrom"], info["lookat"], (0, 1, 0), info["field_of_view"], 1.5, 0.0, focus)
r = rt.Renderer.new_with_rng(cam, scene, world.background(), rt.RenderingParams(50, 266, 400), rt.RecursiveRayTracer(50), rt.SeedableRngator(42))
out = []
for k in range(6):
    r.render_arrays(want_accum=False)
    out.append(round(r.stats["paths"] / r.stats["device_ms"] / 1e3, 1))
print("C1 Mpaths/s per call (first = cold):", out, "layout", r.stats["bvh_layout"])
