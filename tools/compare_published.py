#!/usr/bin/env python
"""RMSE between renders of this repo's CLI (PPM from rt_main, run on a B200) and the reference's own published
renders of the same command lines (/root/reference/*.jpg, seed 42).  The JPEGs are lossy and the render streams
differ (Philox vs PCG64), so this is a statistical check against the REAL reference, complementing the oracle tests.
usage: python tools/compare_published.py   (expects gpurun_out/final_scene_c4.ppm and gpurun_out/random_c2.ppm)"""
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAIRS = [("gpurun_out/final_scene_c4.ppm", "/root/reference/final_scene.jpg", "assets/final_scene_b200_10000spp.jpg",
          "rt_main --world=final_scene --seed=42 --aspect_ratio=1:1 --image_width=800 --samples_per_pixel=10000   (README.md:31-36 of the reference)"),
         ("gpurun_out/random_c2.ppm", "/root/reference/sample_blur.jpg", "assets/random_blur_b200_500spp.jpg",
          "rt_main --world=random --seed=42 --aspect_ratio=3:2 --image_width=1200 --samples_per_pixel=500 --aperture=0.1 --focus_dist=10.0   (README.md:20-25)")]


def read_p3(path):
    toks = open(path).read().split()
    assert toks[0] == "P3" and toks[3] == "255"
    w, h = int(toks[1]), int(toks[2])
    return np.array(toks[4:], dtype=np.uint8).reshape(h, w, 3)


for ppm, ref, out, cmd in PAIRS:
    a = read_p3(os.path.join(ROOT, ppm))
    Image.fromarray(a).save(os.path.join(ROOT, out), quality=92)
    if not os.path.exists(ref):
        print(f"{ppm}: reference render {ref} not available here")
        continue
    r = Image.open(ref).convert("RGB")
    if r.size != (a.shape[1], a.shape[0]):
        r = r.resize((a.shape[1], a.shape[0]), Image.LANCZOS)
    d = (a.astype(np.float64) - np.asarray(r, dtype=np.float64)) / 255.0
    print(cmd)
    print(f"  {a.shape[1]}x{a.shape[0]}: RMSE vs {os.path.basename(ref)} = {np.sqrt((d ** 2).mean()):.5f} (display values in [0,1]); "
          f"mean difference per channel = {np.round(d.reshape(-1, 3).mean(axis=0), 5).tolist()}")
