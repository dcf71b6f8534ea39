"""The two things the persistent kernel decides per scene on the host side (flatten.cpp): which media are "clear" (nothing but
the medium inside a convex boundary, so that a path can hop from event to event without a surface search: wf_chain_step)
and which features the device code must carry (F_* bits: the kernel instance launched holds no code for the others).
CPU part: the flattener's answers and the equivalence of the chain step with the ordinary extend + shade path, on the host
build of the device header.  GPU part: the specialised instances against the generic one."""
import numpy as np
import pytest

import support as S
from mu_lambda_raytracer_b200 import abi
from mu_lambda_raytracer_b200 import renderer as rt

# rt_types.h
F_BOX, F_INSTBOX, F_INSTANCE, F_MOVING, F_BIG, F_MEDIA, F_BOXMEDIA, F_CHECKER, F_NOISE, F_IMAGE, F_SPHERE = (1 << k for k in range(11))


def _world_scene(name):
    world = rt.World(name)
    desc = world.build(42)
    return world, desc, S.EmulScene(desc.ptr)


@pytest.mark.parametrize("name,want", [
    ("random", F_SPHERE | F_BIG),
    ("random_chk", F_SPHERE | F_BIG | F_CHECKER),
    ("cornell_box", F_BOX | F_INSTBOX | F_INSTANCE),
    ("cornell_smoke", F_BOX | F_MEDIA | F_BOXMEDIA),
    ("earth", F_SPHERE | F_IMAGE),
    ("final_scene", F_SPHERE | F_BOX | F_INSTANCE | F_BIG | F_MEDIA | F_NOISE | F_IMAGE),
])
def test_scene_feature_bits(name, want):
    """what worlds.rs puts into each world (worlds.rs:29-484) is what the flattener reports — in particular final_scene has no
    instanced BOX (the rotated group holds spheres, worlds.rs:441-452) and cornell_smoke's boxes exist only as media boundaries"""
    _, _, es = _world_scene(name)
    assert es.features == want, (name, bin(es.features), bin(want))


def test_clear_media_of_the_shipped_worlds():
    """final_scene: the blue ball (medium 0; its glass shell is the same sphere) is clear, the fog around everything is not;
    cornell_smoke: both smoke boxes; a world without media: none"""
    assert _world_scene("final_scene")[2].clear_media == 0b01
    assert _world_scene("cornell_smoke")[2].clear_media == 0b11
    assert _world_scene("random")[2].clear_media == 0


def _medium_scene(intruder):
    b = S.DescBuilder()
    glass = b.material(abi.RT_MAT_DIELECTRIC, ior=1.5)
    grey = b.lambertian(b.solid(0.5, 0.5, 0.5))
    shell = b.sphere((0, 0, 0), 2.0, glass)
    items = [shell, b.medium(b.sphere((0, 0, 0), 2.0, glass), 0.5, (0.2, 0.4, 0.9)), b.sphere((0, -102, 0), 100.0, grey)]
    if intruder == "inside":
        items.append(b.sphere((0.5, 0.2, 0), 0.3, grey))
    elif intruder == "poking":
        items.append(b.sphere((2.2, 0, 0), 0.5, grey))
    elif intruder == "near":  # its box overlaps the boundary's box, the sphere itself stays outside the ball
        items.append(b.sphere((1.8, 1.8, 0), 0.5, grey))
    elif intruder == "smaller shell":
        items.append(b.sphere((0, 0, 0), 1.5, glass))
    return b, b.finish(b.group(abi.RT_NODE_LIST, items), background=abi.RT_BG_GRADIENT)


@pytest.mark.parametrize("intruder,clear", [(None, True), ("inside", False), ("poking", False), ("near", True), ("smaller shell", False)])
def test_clear_medium_rule(intruder, clear):
    """a surface that reaches into the ball disqualifies the medium; the coincident shell and a neighbour outside do not"""
    b, desc = _medium_scene(intruder)
    es = S.EmulScene(desc)
    assert (es.clear_media == 1) == clear


@pytest.mark.parametrize("name,aspect", [("final_scene", 1.0), ("cornell_smoke", 1.0)])
def test_chain_step_equals_extend_and_shade(name, aspect):
    """paths advanced by wf_chain_step inside a clear medium give the same sums, BIT FOR BIT, and the same ray count as paths that
    search the surfaces and go through wf_shade for every segment (same Philox counters, same arithmetic)"""
    world, desc, es = _world_scene(name)
    info = world.camera()
    cam = S.make_camera(info["lookfrom"], info["lookat"], info["field_of_view"], aspect)
    W, spp = 64, 12
    H = int(W / aspect)
    plain, rays0, steps0 = es.render_slots(cam, W, H, spp, chain=False)
    chain, rays1, steps1 = es.render_slots(cam, W, H, spp, chain=True)
    assert steps0 == 0 and steps1 > 0.1 * rays1
    assert rays0 == rays1
    assert np.array_equal(plain, chain)
    mega, _ = es.render(cam, W, H, spp, fixed=True)  # and both are the megakernel routine's paths
    assert np.array_equal(plain, mega)


def test_chain_step_with_an_intruder_is_not_taken():
    b, desc = _medium_scene("inside")
    es = S.EmulScene(desc)
    cam = S.make_camera((0, 1, -8), (0, 0, 0), 40.0, 1.0)
    a, rays0, steps0 = es.render_slots(cam, 48, 48, 8, chain=True)
    assert steps0 == 0
    b2, desc2 = _medium_scene(None)
    es2 = S.EmulScene(desc2)
    plain, r0, _ = es2.render_slots(cam, 48, 48, 8, chain=False)
    chain, r1, steps = es2.render_slots(cam, 48, 48, 8, chain=True)
    assert steps > 0 and r0 == r1 and np.array_equal(plain, chain)


# ------------------------------------------------------------------------------------------------------------------ GPU
def _render(name, aspect, W, spp, seed=11):
    world = rt.World(name)
    scene = rt.Scene(world.build(42))
    info = world.camera()
    cam = S.make_camera(info["lookfrom"], info["lookat"], info["field_of_view"], aspect)
    r = rt.Renderer.new_with_rng(cam, scene, world.background(), rt.RenderingParams(spp, int(W / aspect), W), rt.RecursiveRayTracer(50), rt.SeedableRngator(seed))
    rgb, accum = r.render_arrays()
    stats = dict(r.stats)
    scene.close()
    return rgb, accum.astype(np.float64), stats


@pytest.mark.gpu
@pytest.mark.parametrize("name,aspect", [("random", 1.5), ("cornell_smoke", 1.0), ("final_scene", 1.0), ("earth", 1.5), ("two_spheres", 1.5)])
def test_feature_instances_trace_the_generic_instance_paths(name, aspect, monkeypatch):
    """a kernel instance specialised to the scene's features draws the same Philox numbers and runs the same functions as the
    generic one; the compiler contracts some multiply-adds differently in the two kernels, so a handful of grazing paths
    take another branch — nearly every pixel agrees to float rounding, the ray counts to 1e-3 (as between the pipelines)"""
    W, spp = 128, 32
    monkeypatch.setenv("RT_PS_FEAT", "0")
    rgb0, a0, st0 = _render(name, aspect, W, spp)
    monkeypatch.setenv("RT_PS_FEAT", "1")
    rgb1, a1, st1 = _render(name, aspect, W, spp)
    assert st0["paths"] == st1["paths"]
    assert abs(st0["rays"] - st1["rays"]) <= 1e-3 * st0["rays"]
    close = np.isclose(a0, a1, rtol=1e-4, atol=1e-4 * a0.mean()).all(axis=2)
    assert close.mean() > 0.99, close.mean()
    assert abs(a0.mean() - a1.mean()) < 1e-3 * a0.mean()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["final_scene", "cornell_smoke"])
def test_chain_instance_on_device(name, monkeypatch):
    """the opt-in kernel instance with the chain phase (RT_PS_CHAIN=1) against the generic instance: same paths, same ray count"""
    W, spp = 128, 32
    monkeypatch.setenv("RT_PS_FEAT", "0")
    rgb0, a0, st0 = _render(name, 1.0, W, spp)
    monkeypatch.setenv("RT_PS_CHAIN", "1")
    rgb1, a1, st1 = _render(name, 1.0, W, spp)
    assert st0["paths"] == st1["paths"]
    assert abs(st0["rays"] - st1["rays"]) <= 1e-3 * st0["rays"]
    close = np.isclose(a0, a1, rtol=1e-4, atol=1e-4 * a0.mean()).all(axis=2)
    assert close.mean() > 0.99, close.mean()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["final_scene", "cornell_smoke"])
@pytest.mark.parametrize("log2_paths,log2_cap", [(27, 24), (14, 8)])
def test_chain_queues_on_device(name, log2_paths, log2_cap, monkeypatch):
    """the opt-in second kernel for the paths inside clear media (RT_PS_CHAINQ=1: chain queue -> chain_kernel -> exit queue ->
    next launch, drain rounds at the end) renders what the single kernel renders — also with launches of 2^14 paths and queues
    of 256 records, where most pushes find the queue full and the paths stay in the kernel"""
    W, spp = 96, 24
    rgb0, a0, st0 = _render(name, 1.0, W, spp)
    monkeypatch.setenv("RT_PS_CHAINQ", "1")
    monkeypatch.setenv("RT_PS_CHAINQ_LOG2_PATHS", str(log2_paths))
    monkeypatch.setenv("RT_PS_CHAINQ_LOG2_CAP", str(log2_cap))
    rgb1, a1, st1 = _render(name, 1.0, W, spp)
    assert st0["paths"] == st1["paths"]
    assert abs(st0["rays"] - st1["rays"]) <= 1e-4 * st0["rays"]
    assert st1["kernel_launches"] > st0["kernel_launches"]
    close = np.isclose(a0, a1, rtol=1e-5, atol=1e-5 * a0.mean()).all(axis=2)
    assert close.mean() > 0.999, close.mean()


def test_feature_bits_of_ad_hoc_descriptions():
    """every kind of node / texture sets its bit — a scene outside every specialised mask falls back to the generic kernel instance"""
    b = S.DescBuilder()
    grey = b.lambertian(b.solid(0.5, 0.5, 0.5))
    chk = b.lambertian(b.checker(b.solid(0, 0, 0), b.solid(1, 1, 1)))
    items = [b.moving_sphere((0, 0, 0), (0, 1, 0), 0.5, grey), b.sphere((3, 0, 0), 150.0, chk),
             b.translate((1, 2, 3), b.rotate(1, 30.0, b.block((0, 0, 0), (1, 1, 1), grey))), b.rect(abi.RT_NODE_XZRECT, 0, 1, 0, 1, 5.0, grey),
             b.medium(b.block((10, 10, 10), (12, 12, 12), grey), 0.1, (1, 1, 1))]
    es = S.EmulScene(b.finish(b.group(abi.RT_NODE_LIST, items)))
    assert es.features == F_SPHERE | F_MOVING | F_BIG | F_CHECKER | F_BOX | F_INSTBOX | F_INSTANCE | F_MEDIA | F_BOXMEDIA
    assert es.clear_media == 0  # conservative: the box of the big sphere AROUND the smoke block overlaps the block's (its surface does not reach it)
