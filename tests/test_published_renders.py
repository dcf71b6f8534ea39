"""The only evidence about the REAL reference's output that exists here: the renders it publishes for seed 42
(README.md:20-38 of the reference; decoded losslessly into tests/golden/published_*.png by tests/golden/make_golden.py).
The Rust binary cannot be run (no toolchain), its render streams are PCG64 and the device's are Philox, and the files went
through JPEG — so these are statistical comparisons with stated thresholds:

* CPU suite: the oracle's render of the image region that the 1000 foam spheres of final_scene project into matches the
  published render there for seed 42 and does NOT for another world seed.  The foam is the last thing the world RNG
  places (after 12 553 - 4 324 draws incl. every rejection-sensitive usize draw), so this pins the restated RNG chain
  against the reference's own picture (SURVEY App. A.4's "foam silhouette" check, automated).
* GPU suite: this repo's renders of the reference's README command lines (C4 in full: 800x800, 10 000 spp; C2: 1200x800,
  500 spp, aperture 0.1) against the published pictures: RMSE and mean differences in display units.
"""
import json
import os

import numpy as np
import pytest
from PIL import Image

import mu_lambda_raytracer_b200 as rt
from mu_lambda_raytracer_b200 import abi
import support as S

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def published(name):
    """display values in [0, 1], row 0 = BOTTOM row like the renderer's output"""
    return np.asarray(Image.open(os.path.join(GOLDEN, name)).convert("RGB"), dtype=np.float64)[::-1] / 255.0


def foam_box():
    with open(os.path.join(GOLDEN, "published_render_probe.json")) as f:
        x0, x1, y0, y1 = json.load(f)["final_scene_jpg"]["foam_box_x0_x1_y0_y1"]
    pad = 36
    return max(x0 - pad, 0), min(x1 + pad, 800), max(799 - y1 - pad, 0), min(799 - y0 + pad, 800)  # x0, x1, row0, row1 (rows from the bottom)


def block_mean(a, k=4):
    h, w = a.shape[:2]
    return a[:h // k * k, :w // k * k].reshape(h // k, k, w // k, k, -1).mean(axis=(1, 3))


def region_rmse(radiance_mean, pub, box, k=4):
    """RMSE in display units between k x k block means (taken on LINEAR values, then gamma 2: no Jensen bias at low spp)"""
    x0, x1, r0, r1 = box
    a = np.sqrt(np.clip(block_mean(radiance_mean[r0:r1, x0:x1], k), 0, 1))
    b = np.sqrt(block_mean(pub[r0:r1, x0:x1] ** 2, k))
    return float(np.sqrt(np.mean((a - b) ** 2)))


def test_oracle_foam_region_matches_the_published_final_scene():
    pub = published("published_final_scene.png")
    box = foam_box()
    res = {}
    for seed, spp in ((42, 64), (43, 24)):
        ow = S.OracleWorld("final_scene", seed)
        cam = S.make_camera(ow.lookfrom, ow.lookat, ow.vfov, 1.0)
        a, _, _, _ = ow.render(cam.c, 800, 800, spp, render_seed=5, rows=(box[2], box[3]))
        res[seed] = region_rmse(a / spp, pub, box)
    # measured: 0.049 (seed 42, 64 spp: Monte-Carlo noise of 4x4 block means + JPEG) against 0.18 (other seeds)
    assert res[42] <= 0.07 and res[43] >= 2.5 * res[42], res


def gpu_render(world_name, seed, W, H, spp, aspect, aperture=0.0, focus=None):
    world = rt.World(world_name)
    scene = rt.Scene(world.build(seed))
    info = world.camera()
    if focus is None:
        focus = float(np.linalg.norm(np.asarray(info["lookat"]) - np.asarray(info["lookfrom"])))  # main.rs:190-193
    cam = rt.Camera(info["lookfrom"], info["lookat"], (0, 1, 0), info["field_of_view"], aspect, aperture, focus)
    r = rt.Renderer.new_with_rng(cam, scene, world.background(), rt.RenderingParams(spp, H, W), rt.RecursiveRayTracer(50), rt.SeedableRngator(seed))
    rgb, accum = r.render_arrays()
    scene.close()
    return rgb, accum.astype(np.float64) / spp, r.stats


@pytest.mark.gpu
def test_c4_final_scene_10000spp_matches_the_published_render():
    """--world=final_scene --seed=42 --aspect_ratio=1:1 --image_width=800 --samples_per_pixel=10000 (README.md:31-36)"""
    pub = published("published_final_scene.png")
    rgb, mean, stats = gpu_render("final_scene", 42, 800, 800, 10000, 1.0)
    assert stats["paths"] == 6_400_000_000
    d = rgb.astype(np.float64) / 255.0 - pub
    rmse, bias = float(np.sqrt(np.mean(d ** 2))), np.abs(d.reshape(-1, 3).mean(axis=0))
    # round 1 measured 0.0184 / 0.0007: JPEG quantisation of a high-contrast picture; a wrong scene or radiometry gives > 0.05
    assert rmse <= 0.02 and bias.max() <= 0.002, (rmse, bias)
    box = foam_box()
    here = region_rmse(mean, pub, box)
    _, mean43, _ = gpu_render("final_scene", 43, 800, 800, 500, 1.0)
    other = region_rmse(mean43, pub, box)
    assert here <= 0.03 and other >= 4 * here, (here, other)


@pytest.mark.gpu
def test_c2_random_500spp_defocus_matches_the_published_render():
    """--world=random --seed=42 --aspect_ratio=3:2 --image_width=1200 --samples_per_pixel=500 --aperture=0.1 --focus_dist=10.0 (README.md:20-25)"""
    pub = published("published_sample_blur.png")
    assert pub.shape == (800, 1200, 3)
    rgb, _, stats = gpu_render("random", 42, 1200, 800, 500, 1.5, aperture=0.1, focus=10.0)
    assert stats["paths"] == 480_000_000
    d = rgb.astype(np.float64) / 255.0 - pub
    rmse, bias = float(np.sqrt(np.mean(d ** 2))), np.abs(d.reshape(-1, 3).mean(axis=0))
    assert rmse <= 0.008 and bias.max() <= 0.002, (rmse, bias)  # round 1 measured 0.0065 / 0.0004
