"""EXTENSION with no reference behaviour (BASELINE.json names it, the reference's Ray has no time,
src/vec.rs:215-219): moving spheres and a camera shutter, as in "Ray Tracing: The Next Week".  Parity is therefore
against the oracle's restatement of the BOOK (oracle/scene.hpp MovingSphere, render.hpp Camera.time0/time1), at the same
three levels as everything else: hits at fixed ray times, images within the oracle's own noise floor, and — the part that
touches the reference — scenes WITHOUT motion render exactly as before (tests/test_gpu_api.py bit reproducibility, the
published-render tests)."""
import ctypes as C

import numpy as np
import pytest

import mu_lambda_raytracer_b200 as rt
from mu_lambda_raytracer_b200 import abi
import parity_cases as PC
import support as S
import test_gpu_parity as G


def bouncing_spheres_desc(seed=5):
    """the book's cover scene in small: a ground sphere, a grid of small spheres of which the diffuse ones bounce
    (centre moves up by U[0, 0.5) over the shutter), three big ones; plus one moving sphere under a rotate+translate"""
    b = S.DescBuilder()
    rng = np.random.default_rng(seed)
    items = [b.sphere((0, -1000, 0), 1000.0, b.lambertian(b.solid(0.5, 0.5, 0.5)))]
    for a in range(-5, 5):
        for c in range(-5, 5):
            centre = (a + 0.9 * rng.random(), 0.2, c + 0.9 * rng.random())
            choose = rng.random()
            if choose < 0.7:
                mat = b.lambertian(b.solid(*(rng.random(3) * rng.random(3))))
                items.append(b.moving_sphere(centre, (centre[0], 0.2 + 0.5 * rng.random(), centre[2]), 0.2, mat))
            elif choose < 0.9:
                items.append(b.sphere(centre, 0.2, b.material(abi.RT_MAT_METAL, albedo=tuple(rng.uniform(0.5, 1, 3)), fuzz=0.2)))
            else:
                items.append(b.sphere(centre, 0.2, b.material(abi.RT_MAT_DIELECTRIC, ior=1.5)))
    items.append(b.sphere((0, 1, 0), 1.0, b.material(abi.RT_MAT_DIELECTRIC, ior=1.5)))
    items.append(b.sphere((-4, 1, 0), 1.0, b.lambertian(b.solid(0.4, 0.2, 0.1))))
    items.append(b.translate((2.0, 0.0, 3.0), b.rotate(1, 30.0, b.moving_sphere((0, 0.5, 0), (1.5, 0.5, 0), 0.5, b.lambertian(b.solid(0.8, 0.2, 0.2))))))
    return b, b.finish(b.group(abi.RT_NODE_BVH, items), background=abi.RT_BG_GRADIENT)


RAY_BOX = ([-9, 0.05, -9], [9, 4, 9], [0, 0.4, 0], [5, 0.6, 5])


def hits_at_times(intersect, n):
    b, desc = bouncing_spheres_desc()
    ow = S.OracleWorld(desc=desc)
    rng = np.random.default_rng(8)
    lo, hi, aim, spread = RAY_BOX
    for time in (0.0, 0.37, 1.0):
        rays = S.random_rays(n, rng, lo, hi, target=aim, spread=spread)
        o = ow.hit(rays, time=time)
        for layout in (2, 4):
            g = intersect(desc, rays, layout, time)
            G.check_hits(g, o, rays)
        moved = ow.hit(rays, time=0.0)
        assert (moved["node"] != o["node"]).mean() > 0.005 or time == 0.0  # the motion is visible to the rays


def test_moving_sphere_hits_match_oracle_emulated():
    def intersect(desc, rays, layout, time):
        return S.EmulScene(desc).intersect(rays, mode=0 if layout == 2 else 3, time=time)
    hits_at_times(intersect, 40_000)


@pytest.mark.gpu
def test_moving_sphere_hits_match_oracle_on_device():
    def intersect(desc, rays, layout, time):
        scene = rt.Scene(rt.SceneDescription(desc, owned=False))
        r32 = np.ascontiguousarray(rays, dtype=np.float32)
        hits = (abi.RtHit * len(rays))()
        abi.check(abi.load().rt_intersect_batch_at(scene.handle, -1 if layout == 2 else -2, time, r32.ctypes.data, len(rays), hits))
        scene.close()
        return np.ctypeslib.as_array(hits).copy()
    hits_at_times(intersect, 500_000)


def image_check(render, W, H, spp):
    b, desc = bouncing_spheres_desc()
    ow = S.OracleWorld(desc=desc)
    cam = S.make_camera((13, 2, 3), (0, 0, 0), 20.0, 16 / 9, aperture=0.1, focus_dist=10.0, time0=0.0, time1=1.0)
    g = render(desc, cam).astype(np.float64)
    a1, _, _, _ = ow.render(cam.c, W, H, spp, render_seed=1)
    a2, _, _, _ = ow.render(cam.c, W, H, spp, render_seed=2)
    disp = lambda a: np.sqrt(np.clip(a / spp, 0, 1))
    rm = lambda x, y: float(np.sqrt(np.mean((x - y) ** 2)))
    floor = rm(disp(a1), disp(a2))
    got = 0.5 * (rm(disp(g), disp(a1)) + rm(disp(g), disp(a2)))
    assert got <= 1.1 * floor, (got, floor)
    for c in range(3):
        m = 0.5 * (a1[..., c].mean() + a2[..., c].mean())
        assert abs(g[..., c].mean() - m) <= 0.01 * m + 4 * abs(a1[..., c].mean() - a2[..., c].mean())
    # the blur is really there: with the shutter closed at time 0 the oracle's picture differs by far more than noise
    still = S.make_camera((13, 2, 3), (0, 0, 0), 20.0, 16 / 9, aperture=0.1, focus_dist=10.0)
    a0, _, _, _ = ow.render(still.c, W, H, spp, render_seed=1)
    assert rm(disp(a0), disp(a1)) > 1.3 * floor


def test_motion_blur_image_emulated():
    W, H, spp = 128, 72, 32
    image_check(lambda desc, cam: S.EmulScene(desc).render(cam, W, H, spp, seed=4)[0], W, H, spp)


@pytest.mark.gpu
@pytest.mark.parametrize("pipeline", [abi.RT_PIPELINE_PERSISTENT, abi.RT_PIPELINE_MEGAKERNEL, abi.RT_PIPELINE_WAVEFRONT])
def test_motion_blur_image_on_device(pipeline):
    W, H, spp = 400, 225, 64

    def render(desc, cam):
        scene = rt.Scene(rt.SceneDescription(desc, owned=False))
        r = rt.Renderer.new_with_rng(cam, scene, rt.GradientBackground(), rt.RenderingParams(spp, H, W), rt.RecursiveRayTracer(50), rt.SeedableRngator(4))
        r.pipeline = pipeline
        _, accum = r.render_arrays()
        scene.close()
        return accum
    image_check(render, W, H, spp)
