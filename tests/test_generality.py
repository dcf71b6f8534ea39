"""What the reference's types allow and round 1 refused (RT_ERR_UNSUPPORTED): nested Checker<Odd, Even>
(src/textures.rs:28-49), more than four constant media, deep/skewed BVHs.  (Compound medium boundaries and unordered
Block corners are parity cases of tests/parity_cases.py.)  Each check runs on the host build of the device header
against the oracle here, and again through the C ABI on the GPU (-m gpu)."""
import ctypes as C

import numpy as np
import pytest

import mu_lambda_raytracer_b200 as rt
from mu_lambda_raytracer_b200 import abi
import support as S


def nested_checker_desc():
    b = S.DescBuilder()
    inner = b.checker(b.solid(0.9, 0.1, 0.1), b.solid(0.1, 0.9, 0.1))
    inner2 = b.checker(b.solid(0.1, 0.1, 0.9), inner)
    outer = b.checker(inner, inner2)  # three levels: Checker<Checker<..>, Checker<.., Checker<..>>>
    node = b.sphere((0, 0, 0), 1.0, b.lambertian(outer))
    return b, b.finish(node), outer


def texture_points(n, seed=3):
    rng = np.random.default_rng(seed)
    uvp = np.zeros((n, 5))
    uvp[:, 0:2] = rng.uniform(0, 1, (n, 2))
    uvp[:, 2:5] = rng.uniform(-30, 30, (n, 3))
    return uvp.astype(np.float32).astype(np.float64)


def oracle_texture(desc, tex, uvp):
    ow = S.OracleWorld(desc=desc)
    want = np.zeros((len(uvp), 3))
    assert S.oracle().orc_texture_value(ow.h, tex, uvp.ctypes.data, len(uvp), want.ctypes.data) == 0
    return want


def test_nested_checker_matches_oracle_emulated():
    b, desc, tex = nested_checker_desc()
    uvp = texture_points(50_000)
    es = S.EmulScene(desc)
    got = np.zeros((len(uvp), 3), np.float32)
    u32 = uvp.astype(np.float32)
    S.emul().emul_texture_value_batch(es.h, tex, u32.ctypes.data, len(uvp), got.ctypes.data)
    want = oracle_texture(desc, tex, uvp)
    bad = np.abs(got - want).max(axis=1) > 1e-6  # only points within f32 rounding of a zero of the sines
    assert bad.mean() < 0.002
    # nested checkers see the same p, hence the same side at every level: only the odd-most and the even-most leaf show
    assert len(np.unique(np.round(want, 3), axis=0)) == 2


def test_checker_nesting_too_deep_is_refused():
    b = S.DescBuilder()
    t = b.solid(1, 1, 1)
    for _ in range(10):
        t = b.checker(t, b.solid(0, 0, 0))
    desc = b.finish(b.sphere((0, 0, 0), 1.0, b.lambertian(t)))
    with pytest.raises(RuntimeError, match="nest deeper"):
        S.EmulScene(desc)


def many_media_desc(n_media=6):
    """a row of smoke balls of different densities and colours over a lit floor, under a sky"""
    b = S.DescBuilder()
    items = [b.rect(abi.RT_NODE_XZRECT, -60, 60, -30, 30, 0.0, b.lambertian(b.solid(0.6, 0.6, 0.6)))]
    glass = b.material(abi.RT_MAT_DIELECTRIC, ior=1.5)
    rng = np.random.default_rng(4)
    for k in range(n_media):
        c = (-40.0 + 16.0 * k, 8.0, 0.0)
        items.append(b.medium(b.sphere(c, 7.0, glass), 0.05 + 0.1 * k, tuple(rng.uniform(0.2, 0.95, 3))))
    return b, b.finish(b.group(abi.RT_NODE_LIST, items), background=abi.RT_BG_GRADIENT)


def test_more_than_four_media_emulated():
    b, desc = many_media_desc(6)
    es = S.EmulScene(desc)
    assert es.n_media == 6
    ow = S.OracleWorld(desc=desc)
    cam = S.make_camera((0, 12, -90), (0, 8, 0), 40.0, 2.0)
    W, H, spp = 96, 48, 48
    g, _ = es.render(cam, W, H, spp, seed=9)
    a1, _, _, _ = ow.render(cam.c, W, H, spp, render_seed=1)
    a2, _, _, _ = ow.render(cam.c, W, H, spp, render_seed=2)
    disp = lambda a: np.sqrt(np.clip(np.asarray(a, np.float64) / spp, 0, 1))
    rm = lambda x, y: float(np.sqrt(np.mean((x - y) ** 2)))
    floor = rm(disp(a1), disp(a2))
    got = 0.5 * (rm(disp(g), disp(a1)) + rm(disp(g), disp(a2)))
    assert got <= 1.1 * floor, (got, floor)
    # every ball shows: the columns through the last two balls (media 4 and 5, second Philox block) are tinted as the oracle's
    for k in (4, 5):
        x = int(W * (0.5 - (-40.0 + 16.0 * k) / 65.5 * 0.5))  # mirrored: +x is to the left from this camera
        col = slice(max(x - 3, 0), x + 4)
        assert abs(disp(g)[18:30, col].mean() - 0.5 * (disp(a1)[18:30, col].mean() + disp(a2)[18:30, col].mean())) < 0.03


def concentric_spheres_desc(n=1000):
    b = S.DescBuilder()
    glass = b.material(abi.RT_MAT_DIELECTRIC, ior=1.5)
    items = [b.sphere((0.0, 0.0, 0.0), 1.0 + 0.01 * k, glass) for k in range(n)]
    return b, b.finish(b.group(abi.RT_NODE_BVH, items))


def test_skewed_bvh_stays_within_the_traversal_stack():
    """nested primitives make the SAH tree a comb; the builder must bound its depth (median splits below
    depth 20) instead of overrunning the 48-entry stack, and both tree layouts must still find every hit"""
    b, desc = concentric_spheres_desc(1000)
    es = S.EmulScene(desc)
    assert es.n_prims == 1000 and es.depth <= 46
    rng = np.random.default_rng(2)
    rays = S.random_rays(20_000, rng, [-20, -20, -20], [20, 20, 20], target=[0, 0, 0], spread=[8, 8, 8])
    lin, g2, g4 = es.intersect(rays, mode=1), es.intersect(rays, mode=0), es.intersect(rays, mode=3)
    assert (lin["material"] >= 0).sum() > 5000
    for g in (g2, g4):
        assert np.array_equal(g["material"] >= 0, lin["material"] >= 0)
        hit = lin["material"] >= 0
        assert np.allclose(g["t"][hit], lin["t"][hit], rtol=1e-6, atol=0)


def checker_far_from_origin(run_texture, n):
    """Checker (textures.rs:40-49) at |p| up to 1000, i.e. sin() of up to 5000 rad: the device's f32 product 5 * x is off
    by at most half an ulp of 5000 (2.4e-4 rad), so a point may change side only inside that band around a zero of
    one of the three sines — and nowhere else (an approximate hardware sine would be off by ~1e-3 rad out there)"""
    b = S.DescBuilder()
    tex = b.checker(b.solid(0.2, 0.3, 0.1), b.solid(0.9, 0.9, 0.9))
    desc = b.finish(b.sphere((0, 0, 0), 1.0, b.lambertian(tex)))
    rng = np.random.default_rng(31)
    uvp = np.zeros((n, 5))
    uvp[:, 2:5] = rng.uniform(-1000, 1000, (n, 3))
    uvp = uvp.astype(np.float32).astype(np.float64)
    got = run_texture(desc, tex, uvp)
    want = oracle_texture(desc, tex, uvp)
    bad = np.abs(got - want).max(axis=1) > 1e-6
    band = np.abs(np.sin(5.0 * uvp[:, 2:5])).min(axis=1)
    assert bad.mean() < 1e-3, bad.mean()
    assert np.all(band[bad] < 5e-4), band[bad].max()


def test_checker_far_from_origin_emulated():
    def run(desc, tex, uvp):
        es = S.EmulScene(desc)
        got = np.zeros((len(uvp), 3), np.float32)
        u32 = uvp.astype(np.float32)
        S.emul().emul_texture_value_batch(es.h, tex, u32.ctypes.data, len(uvp), got.ctypes.data)
        return got
    checker_far_from_origin(run, 100_000)


# ------------------------------------------------------------------ the same through the C ABI on the GPU
@pytest.mark.gpu
def test_checker_far_from_origin_on_device():
    def run(desc, tex, uvp):
        scene = rt.Scene(rt.SceneDescription(desc, owned=False))
        got = np.zeros((len(uvp), 3), np.float32)
        u32 = uvp.astype(np.float32)
        abi.check(abi.load().rt_texture_value_batch(scene.handle, tex, u32.ctypes.data, len(uvp), got.ctypes.data))
        scene.close()
        return got
    checker_far_from_origin(run, 1_000_000)



@pytest.mark.gpu
def test_nested_checker_matches_oracle_on_device():
    b, desc, tex = nested_checker_desc()
    scene = rt.Scene(rt.SceneDescription(desc, owned=False))
    uvp = texture_points(200_000)
    got = np.zeros((len(uvp), 3), np.float32)
    u32 = uvp.astype(np.float32)
    abi.check(abi.load().rt_texture_value_batch(scene.handle, tex, u32.ctypes.data, len(uvp), got.ctypes.data))
    want = oracle_texture(desc, tex, uvp)
    assert (np.abs(got - want).max(axis=1) > 1e-6).mean() < 0.002
    scene.close()


@pytest.mark.gpu
@pytest.mark.parametrize("pipeline", [abi.RT_PIPELINE_PERSISTENT, abi.RT_PIPELINE_MEGAKERNEL, abi.RT_PIPELINE_WAVEFRONT])
def test_more_than_four_media_on_device(pipeline):
    b, desc = many_media_desc(6)
    scene = rt.Scene(rt.SceneDescription(desc, owned=False))
    assert scene.info()["media"] == 6
    ow = S.OracleWorld(desc=desc)
    cam = S.make_camera((0, 12, -90), (0, 8, 0), 40.0, 2.0)
    W, H, spp = 192, 96, 128
    r = rt.Renderer.new_with_rng(cam, scene, rt.GradientBackground(), rt.RenderingParams(spp, H, W), rt.RecursiveRayTracer(50), rt.SeedableRngator(5))
    r.pipeline = pipeline
    _, g = r.render_arrays()
    a1, _, _, _ = ow.render(cam.c, W, H, spp, render_seed=1)
    a2, _, _, _ = ow.render(cam.c, W, H, spp, render_seed=2)
    disp = lambda a: np.sqrt(np.clip(np.asarray(a, np.float64) / spp, 0, 1))
    rm = lambda x, y: float(np.sqrt(np.mean((x - y) ** 2)))
    floor = rm(disp(a1), disp(a2))
    got = 0.5 * (rm(disp(g), disp(a1)) + rm(disp(g), disp(a2)))
    assert got <= 1.1 * floor, (got, floor)
    for c in range(3):
        m = 0.5 * (a1[..., c].mean() + a2[..., c].mean())
        assert abs(g[..., c].mean() - m) <= 0.01 * m + 4 * abs(a1[..., c].mean() - a2[..., c].mean())
    scene.close()


@pytest.mark.gpu
def test_skewed_bvh_on_device():
    b, desc = concentric_spheres_desc(1000)
    scene = rt.Scene(rt.SceneDescription(desc, owned=False))
    ow = S.OracleWorld(desc=desc)
    rng = np.random.default_rng(2)
    rays = S.random_rays(200_000, rng, [-20, -20, -20], [20, 20, 20], target=[0, 0, 0], spread=[8, 8, 8])
    import test_gpu_parity as G
    o = ow.hit(rays)
    for node in (-1, -2):
        G.check_hits(G.gpu_intersect(scene, rays, node=node), o, rays)
    scene.close()
