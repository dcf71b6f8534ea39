"""Size-independent properties of the path (no oracle image needed):

* white furnace — every surface and medium has albedo 1 and the environment is uniformly white, so the reference's
  estimator (product of attenuations times one terminal term, raytrace.rs:79-101) returns exactly 1 for every path
  that is not cut by the depth limit: the image must be 1 everywhere.  Any lost or double-counted energy in the
  camera, the intersection code, the samplers, the media or the accumulation shows up as a deviation.
* additivity — the samples [0, a) and [a, b) rendered separately add up to the samples [0, b) rendered at once, ray
  for ray (this is what the multi-GPU sample slices rely on); checked at the full C4 size on the GPU.
"""
import ctypes as C

import numpy as np
import pytest

import mu_lambda_raytracer_b200 as rt
from mu_lambda_raytracer_b200 import abi
import support as S


def furnace_desc():
    b = S.DescBuilder()
    white = b.solid(1.0, 1.0, 1.0)
    lam = b.lambertian(white)
    lam_chk = b.lambertian(b.checker(white, b.solid(1.0, 1.0, 1.0)))
    glass = b.material(abi.RT_MAT_DIELECTRIC, ior=1.5)
    iso = b.material(abi.RT_MAT_ISOTROPIC, tex=white)
    items = [b.sphere((0.0, -100.5, -1.0), 100.0, lam_chk), b.sphere((0.0, 0.0, -1.0), 0.5, lam), b.sphere((1.1, 0.0, -1.0), 0.5, glass),
             b.sphere((1.1, 0.0, -1.0), -0.4, glass), b.translate((-1.6, -0.5, -1.4), b.rotate(1, 25.0, b.block((0.0, 0.0, 0.0), (0.7, 0.9, 0.7), lam)))]
    fogs = []
    # three media (overlapping each other and the surfaces): a dense ball, a thin global haze, a rotated box of smoke
    for boundary, density in ((b.sphere((-0.2, 0.3, -0.2), 0.45, glass), 3.0), (b.sphere((0.0, 0.0, 0.0), 50.0, glass), 0.02),
                              (b.translate((0.4, -0.4, 0.2), b.rotate(1, -18.0, b.block((0.0, 0.0, 0.0), (0.6, 0.6, 0.6), lam))), 1.5)):
        fog = b.medium(boundary, density, (1.0, 1.0, 1.0))
        b.materials[b.nodes[fog].material] = b.materials[iso]  # the medium's phase material: white isotropic
        fogs.append(fog)
    root = b.group(abi.RT_NODE_LIST, [b.group(abi.RT_NODE_BVH, items)] + fogs)
    desc = b.finish(root, background=abi.RT_BG_GRADIENT)
    for i in range(3):
        desc.contents.background_top[i] = 1.0
        desc.contents.background_bottom[i] = 1.0
    return b, desc


def check_furnace(accum, spp, what):
    img = accum.astype(np.float64) / spp
    exact = np.abs(img - 1.0) < 2e-6
    # a path cut by max_depth returns 0 (raytrace.rs:87-89): rare here, but it exists (light trapped inside the glass),
    # so a few pixels miss one of their samples; nothing may ever exceed 1
    assert exact.mean() > 0.99, (what, exact.mean(), img.min(), img.max())
    assert img.max() < 1.0 + 2e-6 and img.mean() > 0.999, (what, img.mean(), img.max())


def test_white_furnace_oracle_and_emulated_device_math():
    b, desc = furnace_desc()
    cam = S.make_camera((-2.0, 2.0, 1.0), (0.0, 0.0, -1.0), 55.0, 1.5)
    W, H, spp = 96, 64, 8
    ow = S.OracleWorld(desc=desc)
    a, _, counters, _ = ow.render(cam.c, W, H, spp, render_seed=3)
    check_furnace(a, spp, "oracle")
    es = S.EmulScene(desc)
    e, rays = es.render(cam, W, H, spp, seed=3)
    check_furnace(e, spp, "emulated device math")
    assert abs(rays / (W * H * spp) - counters[1] / counters[0]) < 0.05 * counters[1] / counters[0]


@pytest.mark.gpu
@pytest.mark.parametrize("pipeline", [abi.RT_PIPELINE_PERSISTENT, abi.RT_PIPELINE_WAVEFRONT, abi.RT_PIPELINE_MEGAKERNEL])
def test_white_furnace_on_device(pipeline):
    b, desc = furnace_desc()
    scene = rt.Scene(rt.SceneDescription(desc, owned=False))
    cam = S.make_camera((-2.0, 2.0, 1.0), (0.0, 0.0, -1.0), 55.0, 1.5)
    W, H, spp = 384, 256, 32
    r = rt.Renderer.new_with_rng(cam, scene, rt.GradientBackground(), rt.RenderingParams(spp, H, W), rt.RecursiveRayTracer(50), rt.SeedableRngator(3))
    r.pipeline = pipeline
    _, accum = r.render_arrays()
    check_furnace(accum, spp, f"pipeline {pipeline}")
    scene.close()


def _render_range(scene, cam, W, H, spp_total, begin, count, seed):
    p = abi.RtParams()
    p.width, p.height, p.samples_per_pixel, p.max_depth, p.seed = W, H, spp_total, 50, seed
    p.sample_begin, p.sample_count, p.pipeline, p.device = begin, count, abi.RT_PIPELINE_AUTO, -1
    accum = np.empty((H, W, 3), np.float32)
    st = abi.RtStats()
    abi.check(abi.load().rt_render(scene.handle, C.byref(cam.c), C.byref(p), accum.ctypes.data, None, abi.RtProgressFn(), None, C.byref(st)))
    return accum.astype(np.float64), int(st.rays), int(st.paths)


@pytest.mark.gpu
def test_sample_ranges_add_up_at_full_c4_size():
    world = rt.World("final_scene")
    scene = rt.Scene(world.build(42))
    info = world.camera()
    cam = S.make_camera(info["lookfrom"], info["lookat"], info["field_of_view"], 1.0)
    W = H = 800
    a, rays_a, paths_a = _render_range(scene, cam, W, H, 64, 0, 24, 42)
    b, rays_b, paths_b = _render_range(scene, cam, W, H, 64, 24, 40, 42)
    c, rays_c, paths_c = _render_range(scene, cam, W, H, 64, 0, 64, 42)
    assert paths_a + paths_b == paths_c == W * H * 64
    assert rays_a + rays_b == rays_c  # the same Philox streams -> the same paths, ray for ray
    assert np.allclose(a + b, c, rtol=1e-6, atol=0)  # the sums are exact integers on the device; only the float conversion of each rounds
    # and another seed is another image
    d, _, _ = _render_range(scene, cam, W, H, 64, 0, 64, 43)
    assert not np.allclose(c, d, rtol=1e-2, atol=1e-2 * c.mean())
    scene.close()
