"""EXTENSION: the GPU-side BVH build for large scenes (csrc/rt_lbvh.cu; the reference builds on the host,
src/bhv.rs:122-145).  The tree it produces must find exactly what the host-built tree and a brute-force scan find."""
import ctypes as C
import os
import time

import numpy as np
import pytest

import mu_lambda_raytracer_b200 as rt
from mu_lambda_raytracer_b200 import abi
import support as S
import test_gpu_parity as G

pytestmark = pytest.mark.gpu


class forced:  # RT_BVH_GPU_MIN is read by the library at scene creation
    def __init__(self, value):
        self.value, self.old = str(value), None

    def __enter__(self):
        self.old = os.environ.get("RT_BVH_GPU_MIN")
        os.environ["RT_BVH_GPU_MIN"] = self.value

    def __exit__(self, *a):
        if self.old is None:
            del os.environ["RT_BVH_GPU_MIN"]
        else:
            os.environ["RT_BVH_GPU_MIN"] = self.old


@pytest.mark.parametrize("name", ["final_scene", "random", "cornell_box"])
def test_device_built_tree_matches_oracle_and_host_tree(name):
    world = rt.World(name)
    desc = world.build(42)
    ow = S.OracleWorld(name, 42)
    lo, hi, aim, spread = G.RAY_BOXES[name]
    rays = S.random_rays(400_000, np.random.default_rng(3), lo, hi, target=aim, spread=spread)
    with forced(2):
        dev = rt.Scene(desc)
    host = rt.Scene(desc)
    bi = dev.build_info()
    assert bi["built_on_device"] and bi["build_ms"] > 0 and 2 <= bi["depth"] <= 46 and not host.build_info()["built_on_device"]
    assert dev.info()["bvh_nodes"] == 2 * dev.info()["prims"] - 1
    g = G.gpu_intersect(dev, rays)
    G.check_hits(g, ow.hit(rays), rays)
    h = G.gpu_intersect(host, rays)
    assert np.array_equal(g["material"] >= 0, h["material"] >= 0)
    hit = g["material"] >= 0
    assert np.array_equal(g["t"][hit], h["t"][hit])  # same primitives tested with the same arithmetic: the same distances
    # and the render loop runs on it (binary layout: a device-built scene has no 4-wide tree)
    info = world.camera()
    cam = S.make_camera(info["lookfrom"], info["lookat"], info["field_of_view"], 1.0)
    images = []
    for scene in (dev, host):
        r = rt.Renderer.new_with_rng(cam, scene, world.background(), rt.RenderingParams(16, 96, 96), rt.RecursiveRayTracer(50), rt.SeedableRngator(9))
        r.bvh_layout = 2
        rgb, _ = r.render_arrays()
        images.append(rgb)
        assert r.stats["bvh_layout"] == 2
    assert (images[0] != images[1]).mean() < 0.002  # identical paths except where two primitives tie
    dev.close(), host.close()


def test_large_scene_is_built_on_the_gpu_and_agrees_with_brute_force():
    n = 120_000
    rng = np.random.default_rng(12)
    b = S.DescBuilder()
    mats = [b.lambertian(b.solid(*rng.uniform(0.1, 0.9, 3))) for _ in range(8)]
    centres = rng.uniform(-400, 400, (n, 3))
    radii = rng.uniform(0.5, 4.0, n)
    t0 = time.time()
    inner = b.group(abi.RT_NODE_BVH, [b.sphere(tuple(c), float(r), mats[i % 8]) for i, (c, r) in enumerate(zip(centres, radii))])
    root = b.group(abi.RT_NODE_LIST, [inner])
    desc = b.finish(root, background=abi.RT_BG_GRADIENT)
    t1 = time.time()
    scene = rt.Scene(rt.SceneDescription(desc, owned=False))  # n >= RT_GPU_BUILD_MIN: linear BVH on the GPU
    t2 = time.time()
    bi = scene.build_info()
    assert bi["built_on_device"] and scene.info()["prims"] == n and bi["depth"] <= 46
    with forced(10 ** 9):
        host = rt.Scene(rt.SceneDescription(desc, owned=False))  # the host's SAH sweep over the same primitives
    t3 = time.time()
    print(f"\n{n} spheres: description {t1 - t0:.2f} s; rt_scene_create with the GPU build {t2 - t1:.2f} s (device part {bi['build_ms']:.2f} ms, depth {bi['depth']}); "
          f"with the host SAH build {t3 - t2:.2f} s (depth {host.build_info()['depth']})")
    rays = S.random_rays(4000, rng, [-600, -600, -600], [600, 600, 600], target=[0, 0, 0], spread=[350, 350, 350])
    g = G.gpu_intersect(scene, rays)
    brute = G.gpu_intersect(scene, rays, node=inner)  # a sub-tree query scans its primitives linearly
    h = G.gpu_intersect(host, rays)
    assert (brute["material"] >= 0).sum() > 1000
    for other in (brute, h):
        assert np.array_equal(g["material"] >= 0, other["material"] >= 0)
        hit = g["material"] >= 0
        assert np.array_equal(g["t"][hit], other["t"][hit])
    # throughput of the render loop on both trees (the SAH tree traverses better; the linear one builds ~1000x faster)
    cam = S.make_camera((0, 0, -1500), (0, 0, 0), 40.0, 1.0)
    for label, sc in (("device-built", scene), ("host SAH", host)):
        r = rt.Renderer.new_with_rng(cam, sc, rt.GradientBackground(), rt.RenderingParams(32, 400, 400), rt.RecursiveRayTracer(50), rt.SeedableRngator(1))
        r.render_arrays()
        print(f"  {label}: {r.stats['paths'] / r.stats['device_ms'] / 1e3:.0f} Mpaths/s, {r.stats['rays'] / r.stats['paths']:.2f} rays/path")
    scene.close(), host.close()
