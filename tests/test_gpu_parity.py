"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path, called through the C ABI, against the
oracle on the same seeded inputs.  Three levels as in BASELINE.json north_star / SURVEY §8c."""
import ctypes as C

import numpy as np
import pytest

import mu_lambda_raytracer_b200 as rt
from mu_lambda_raytracer_b200 import abi
import support as S

pytestmark = pytest.mark.gpu

RAY_BOXES = {  # origin box, aim point, aim spread per world
    "simple": ([-3, 0.01, -3], [3, 3, 2], [0, 0, -1], [1.5, 0.5, 0.5]),
    "random": ([-15, 0.01, -15], [15, 5, 15], [0, 0.3, 0], [12, 0.5, 12]),
    "random_chk": ([-15, 0.01, -15], [15, 5, 15], [0, 0.3, 0], [12, 0.5, 12]),
    "two_spheres": ([-15, 0.01, -15], [15, 8, 15], [0, 1, 0], [4, 3, 4]),
    "simple_light": ([-15, 0.01, -15], [15, 10, 15], [0, 2, 0], [5, 5, 5]),
    "cornell_box": ([1, 1, -800], [554, 554, 554], [278, 278, 278], [278, 278, 278]),
    "cornell_smoke": ([1, 1, -800], [554, 554, 554], [278, 278, 278], [278, 278, 278]),
    "earth": ([-15, -15, -15], [15, 15, 15], [0, 0, 0], [2, 2, 2]),
    "debug_perlin": ([0, 0, -600], [600, 600, 300], [278, 278, 0], [80, 80, 80]),
    "final_scene": ([-300, 50, -600], [600, 500, 600], [200, 200, 250], [400, 250, 300]),
}


def gpu_intersect(scene, rays, node=-1):
    r32 = np.ascontiguousarray(rays, dtype=np.float32)
    hits = (abi.RtHit * len(rays))()
    abi.check(abi.load().rt_intersect_batch(scene.handle, node, r32.ctypes.data, len(rays), hits))
    return np.ctypeslib.as_array(hits).copy()


def check_hits(g, o, rays, grazing=0.02, tol=1e-5):
    """hit/miss identical outside a grazing band; position along the ray within tol relative to the scale of the
    coordinates involved; normals within 1e-4 (f32 unit vectors)."""
    ohit, ghit = o["hit"] == 1, g["material"] >= 0
    dlen = np.linalg.norm(rays[:, 3:6], axis=1)
    dirn = rays[:, 3:6] / dlen[:, None]
    cos = np.abs(np.sum(dirn * o["normal"], axis=1))
    cos_g = np.abs(np.sum(dirn * g["normal"].astype(np.float64), axis=1))  # where only the device hit, judge grazing by ITS normal
    mism = ohit != ghit
    # a miss/hit disagreement is tolerated only if the oracle's hit is grazing or sits at the very end of the range
    # ... or passes within f32 rounding of the primitive's rim (u or v at 0 or 1: the edge of a rect / box face)
    rim = lambda h: (np.minimum(h["u"], 1.0 - h["u"]) < 2e-5) | (np.minimum(h["v"], 1.0 - h["v"]) < 2e-5)
    hard = mism & ~(ohit & ((cos < grazing) | rim(o))) & ~(ghit & ~ohit & ((cos_g < grazing) | rim(g)))
    assert hard.sum() <= max(2, len(rays) // 100000), f"{hard.sum()} hit/miss mismatches outside the grazing band"
    both = ohit & ghit & (cos >= grazing)
    scale = np.abs(rays[:, :3]).max(axis=1) + np.abs(o["t"]) * dlen + 1.0
    err = np.abs(g["t"].astype(np.float64) - o["t"]) * dlen
    bad = both & (err > tol * scale)
    assert bad.sum() <= len(rays) // 20000, f"{bad.sum()} hits off by more than {tol} relative (max {np.max(err[both] / scale[both]):.2e})"
    same_prim = both & (g["prim"] == o["node"])
    nerr = np.abs(g["normal"].astype(np.float64) - o["normal"]).max(axis=1)
    assert (same_prim & (nerr > 1e-4)).sum() <= len(rays) // 20000
    assert (both & ~same_prim).sum() <= len(rays) // 500, "closest primitive differs (beyond ties on shared edges)"
    assert np.all(g["material"][same_prim] == o["material"][same_prim])
    ff = same_prim & (g["front_face"] != o["front_face"])
    assert ff.sum() <= len(rays) // 20000


@pytest.mark.parametrize("layout", [2, 4])
@pytest.mark.parametrize("name", list(RAY_BOXES))
def test_closest_hit_matches_oracle(name, layout):
    """whole-world closest hit through the binary (32-byte nodes) and the 4-wide (128-byte nodes) BVH"""
    rng = np.random.default_rng(sum(map(ord, name)) % 1000)
    desc = rt.World(name).build(42)
    scene = rt.Scene(desc)
    ow = S.OracleWorld(name, 42)
    lo, hi, aim, spread = RAY_BOXES[name]
    n = 1_000_000 if name in ("random", "final_scene", "cornell_smoke") else 200_000
    rays = S.random_rays(n, rng, lo, hi, target=aim, spread=spread)
    g = gpu_intersect(scene, rays, node=-1 if layout == 2 else -2)
    o = ow.hit(rays)
    check_hits(g, o, rays)
    scene.close()


PIPELINES = {"megakernel": abi.RT_PIPELINE_MEGAKERNEL, "wavefront": abi.RT_PIPELINE_WAVEFRONT, "persistent": abi.RT_PIPELINE_PERSISTENT}


@pytest.mark.parametrize("pipeline", list(PIPELINES))
@pytest.mark.parametrize("name,aspect,spp", [("random", 1.5, 64), ("cornell_smoke", 1.0, 64), ("final_scene", 1.0, 64)])
def test_image_rmse_within_noise_floor(name, aspect, spp, pipeline):
    world = rt.World(name)
    desc = world.build(42)
    scene = rt.Scene(desc)
    ow = S.OracleWorld(name, 42)
    cam = S.make_camera(ow.lookfrom, ow.lookat, ow.vfov, aspect)
    W = 120
    H = int(W / aspect)
    r = rt.Renderer.new_with_rng(cam, scene, world.background(), rt.RenderingParams(spp, H, W), rt.RecursiveRayTracer(50), rt.SeedableRngator(42))
    r.pipeline = PIPELINES[pipeline]
    rgb, accum = r.render_arrays()
    assert r.stats["pipeline"] == PIPELINES[pipeline]
    a1, rgb1, c1, _ = ow.render(cam.c, W, H, spp, render_seed=42)
    a2, _, _, _ = ow.render(cam.c, W, H, spp, render_seed=977)
    rm = lambda x, y: float(np.sqrt(np.mean((x - y) ** 2)))
    disp = lambda a: np.sqrt(np.clip(a / spp, 0.0, 1.0))  # the reference's display transform before quantisation
    floor = rm(disp(a1), disp(a2))
    got = 0.5 * (rm(disp(accum.astype(np.float64)), disp(a1)) + rm(disp(accum.astype(np.float64)), disp(a2)))
    assert got <= 1.1 * floor, f"RMSE {got:.5f} vs noise floor {floor:.5f}"
    # unbiasedness: mean radiance agrees within a few standard errors of the oracle-vs-oracle difference
    m_g, m_1, m_2 = accum.mean() / spp, a1.mean() / spp, a2.mean() / spp
    assert abs(m_g - 0.5 * (m_1 + m_2)) <= 4 * abs(m_1 - m_2) + 0.01 * m_1
    assert abs(r.stats["rays"] / r.stats["paths"] - c1[1] / c1[0]) < 0.05 * c1[1] / c1[0]
    # the int32 image is exactly to_rgb of the float sums the same call returned
    x = np.sqrt(accum.astype(np.float64) * (1.0 / spp))
    exp = (255.999 * np.clip(x, 0.0, 0.99999999)).astype(np.int32)
    assert np.array_equal(rgb, exp)
    scene.close()


@pytest.mark.parametrize("name,aspect", [("random", 1.5), ("cornell_smoke", 1.0), ("final_scene", 1.0), ("cornell_box", 1.0), ("simple", 16 / 9)])
def test_wavefront_and_megakernel_agree(name, aspect):
    """Both pipelines draw the same Philox numbers for the same (pixel, sample, segment), so they trace the SAME
    paths up to the compiler's different FMA contraction of the shared device functions in the two kernels (a
    handful of grazing paths take another branch): ray counts agree to 1e-3 and nearly all pixels to float rounding."""
    world = rt.World(name)
    desc = world.build(42)
    scene = rt.Scene(desc)
    info = world.camera()
    cam = S.make_camera(info["lookfrom"], info["lookat"], info["field_of_view"], aspect)
    W, spp = 96, 24
    H = int(W / aspect)
    out = {}
    for pname, pid in PIPELINES.items():
        r = rt.Renderer.new_with_rng(cam, scene, world.background(), rt.RenderingParams(spp, H, W), rt.RecursiveRayTracer(50), rt.SeedableRngator(7))
        r.pipeline = pid
        rgb, accum = r.render_arrays()
        out[pname] = (rgb, accum.astype(np.float64), r.stats)
    a = out["megakernel"]
    for other in ("wavefront", "persistent"):
        b = out[other]
        assert a[2]["paths"] == b[2]["paths"] == W * H * spp
        assert abs(a[2]["rays"] - b[2]["rays"]) <= 1e-3 * a[2]["rays"], other
        close = np.isclose(a[1], b[1], rtol=1e-4, atol=1e-4 * a[1].mean()).all(axis=2)
        assert close.mean() > 0.98, (other, close.mean())
        assert abs(a[1].mean() - b[1].mean()) < 2e-3 * a[1].mean(), other
    scene.close()


def test_wavefront_many_rounds_small_pool(monkeypatch):
    """a pool much smaller than the job: thousands of regenerations, graph batches and the termination logic"""
    monkeypatch.setenv("RT_WF_SLOTS", "4096")
    world = rt.World("cornell_smoke")
    desc = world.build(0)
    scene = rt.Scene(desc)
    info = world.camera()
    cam = S.make_camera(info["lookfrom"], info["lookat"], info["field_of_view"], 1.0)
    res = {}
    for pname, pid in PIPELINES.items():
        r = rt.Renderer.new_with_rng(cam, scene, world.background(), rt.RenderingParams(16, 64, 64), rt.RecursiveRayTracer(50), rt.SeedableRngator(3))
        r.pipeline = pid
        _, accum = r.render_arrays()
        res[pname] = (accum.astype(np.float64), r.stats)
    for other in ("wavefront", "persistent"):
        assert abs(res["megakernel"][1]["rays"] - res[other][1]["rays"]) <= 2e-3 * res["megakernel"][1]["rays"]
        close = np.isclose(res["megakernel"][0], res[other][0], rtol=1e-4, atol=1e-3).all(axis=2)
        assert close.mean() > 0.97, (other, close.mean())
    scene.close()


@pytest.mark.parametrize("W,H,spp,depth", [(2, 2, 1, 50), (7, 5, 3, 50), (33, 17, 2, 1), (16, 16, 4, 0), (40, 24, 5, 3)])
def test_edge_sizes_and_depths_all_pipelines(W, H, spp, depth):
    """tiny and ragged images, one sample, max_depth 1 (camera ray only) and 0 (every path is Color::ZERO at once,
    raytrace.rs:87-89): every pipeline traces exactly W*H*spp paths and agrees with the megakernel"""
    world = rt.World("final_scene")
    desc = world.build(42)
    scene = rt.Scene(desc)
    info = world.camera()
    cam = S.make_camera(info["lookfrom"], info["lookat"], info["field_of_view"], W / H)
    out = {}
    for pname, pid in PIPELINES.items():
        r = rt.Renderer.new_with_rng(cam, scene, world.background(), rt.RenderingParams(spp, H, W), rt.RecursiveRayTracer(depth), rt.SeedableRngator(99))
        r.pipeline = pid
        rgb, accum = r.render_arrays()
        assert r.stats["paths"] == W * H * spp
        assert np.isfinite(accum).all() and (accum >= 0).all()
        out[pname] = (accum.astype(np.float64), r.stats["rays"])
    base, base_rays = out["megakernel"]
    if depth == 0:
        for a, rays in out.values():
            assert rays == 0 and not a.any()
    else:
        assert base_rays >= W * H * spp
        for pname, (a, rays) in out.items():
            assert abs(int(rays) - int(base_rays)) <= max(2, 0.01 * base_rays), pname
            close = np.isclose(a, base, rtol=1e-4, atol=1e-4 * max(base.mean(), 1e-6)).all(axis=2)
            assert close.mean() >= 0.95, (pname, close.mean())
    scene.close()


def test_same_seed_same_image_different_seed_different_image():
    world = rt.World("cornell_smoke")
    scene = rt.Scene(world.build(0))
    info = world.camera()
    cam = S.make_camera(info["lookfrom"], info["lookat"], info["field_of_view"], 1.0)

    def render(seed):
        r = rt.Renderer.new_with_rng(cam, scene, world.background(), rt.RenderingParams(8, 48, 48), rt.RecursiveRayTracer(50), rt.SeedableRngator(seed))
        _, accum = r.render_arrays()
        return accum.astype(np.float64), r.stats["rays"]

    a1, r1 = render(5)
    a2, r2 = render(5)
    b, _ = render(6)
    assert r1 == r2  # the same Philox streams: the same paths, float sums differ only by the order of the REDs
    assert np.allclose(a1, a2, rtol=1e-5, atol=1e-5)
    assert not np.allclose(a1, b, rtol=1e-3, atol=1e-3)
    scene.close()


@pytest.mark.parametrize("name,aspect,kw", [("final_scene", 1.0, {}), ("random", 1.5, dict(aperture=0.1, focus_dist=10.0)), ("cornell_smoke", 1.0, {})])
def test_no_bias_at_high_sample_count(name, aspect, kw):
    """The RMSE criterion at 1024 spp, where Monte-Carlo noise is 4x smaller than in the 64-spp test and a bias of the
    f32 device path (self-intersection, culling, sampler) would stand out of it: RMSE(device, oracle) must stay
    within 1.1x the oracle's own seed-to-seed RMSE."""
    world = rt.World(name)
    scene = rt.Scene(world.build(42))
    ow = S.OracleWorld(name, 42)
    cam = S.make_camera(ow.lookfrom, ow.lookat, ow.vfov, aspect, **kw)
    W, spp = 72, 1024
    H = int(W / aspect)
    r = rt.Renderer.new_with_rng(cam, scene, world.background(), rt.RenderingParams(spp, H, W), rt.RecursiveRayTracer(50), rt.SeedableRngator(4242))
    _, accum = r.render_arrays()
    a1, _, _, _ = ow.render(cam.c, W, H, spp, render_seed=1)
    a2, _, _, _ = ow.render(cam.c, W, H, spp, render_seed=2)
    disp = lambda a: np.sqrt(np.clip(a / spp, 0.0, 1.0))
    rm = lambda x, y: float(np.sqrt(np.mean((x - y) ** 2)))
    floor = rm(disp(a1), disp(a2))
    g = disp(accum.astype(np.float64))
    got = 0.5 * (rm(g, disp(a1)) + rm(g, disp(a2)))
    assert got <= 1.1 * floor, f"{name}: RMSE {got:.5f} vs noise floor {floor:.5f} at {spp} spp"
    # linear radiance, channel means: within 1 % of the oracle's
    m_g, m_o = accum.reshape(-1, 3).mean(axis=0) / spp, 0.5 * (a1 + a2).reshape(-1, 3).mean(axis=0) / spp
    assert np.all(np.abs(m_g - m_o) <= 0.01 * m_o + 1e-4), (m_g, m_o)
    scene.close()


ALL_WORLDS = [("simple", 16 / 9), ("random", 1.5), ("random_chk", 1.5), ("two_spheres", 16 / 9), ("simple_light", 16 / 9), ("cornell_box", 1.0),
              ("cornell_smoke", 1.0), ("earth", 16 / 9), ("debug_perlin", 1.0), ("final_scene", 1.0)]


@pytest.mark.parametrize("name,aspect", ALL_WORLDS)
def test_every_world_image_rmse_within_noise_floor(name, aspect):
    """all ten worlds of the registry (worlds.rs:471-484) through the default pipeline: negative-radius glass, checker on
    world-space p, Perlin on large spheres, rect and sphere emitters, rotated boxes as surfaces, the earth texture"""
    world = rt.World(name)
    scene = rt.Scene(world.build(42))
    ow = S.OracleWorld(name, 42)
    cam = S.make_camera(ow.lookfrom, ow.lookat, ow.vfov, aspect)
    W, spp = 80, 96
    H = int(W / aspect)
    r = rt.Renderer.new_with_rng(cam, scene, world.background(), rt.RenderingParams(spp, H, W), rt.RecursiveRayTracer(50), rt.SeedableRngator(7))
    _, accum = r.render_arrays()
    assert r.stats["pipeline"] == abi.RT_PIPELINE_PERSISTENT  # what AUTO resolves to
    a1, _, c1, _ = ow.render(cam.c, W, H, spp, render_seed=11)
    a2, _, _, _ = ow.render(cam.c, W, H, spp, render_seed=12)
    disp = lambda a: np.sqrt(np.clip(a / spp, 0.0, 1.0))
    rm = lambda x, y: float(np.sqrt(np.mean((x - y) ** 2)))
    floor = rm(disp(a1), disp(a2))
    g = disp(accum.astype(np.float64))
    got = 0.5 * (rm(g, disp(a1)) + rm(g, disp(a2)))
    assert got <= 1.1 * floor + 1e-4, f"{name}: RMSE {got:.5f} vs noise floor {floor:.5f}"
    assert abs(r.stats["rays"] / r.stats["paths"] - c1[1] / c1[0]) < 0.05 * c1[1] / c1[0]
    scene.close()


def test_scene_recreation_reuses_cached_device_memory():
    """create / render / destroy in a loop (what a frame-by-frame host such as movie.py does): the blocks of a destroyed
    scene are handed to the next one (csrc/dev_cache.h), results stay identical, other scenes are not disturbed"""
    world = rt.World("final_scene")
    desc = world.build(42)
    info = world.camera()
    cam = S.make_camera(info["lookfrom"], info["lookat"], info["field_of_view"], 1.0)
    other = rt.Scene(rt.World("cornell_smoke").build(0))
    ref = None
    for k in range(5):
        scene = rt.Scene(desc)
        r = rt.Renderer.new_with_rng(cam, scene, world.background(), rt.RenderingParams(8, 64, 64), rt.RecursiveRayTracer(50), rt.SeedableRngator(1))
        _, accum = r.render_arrays()
        rays = r.stats["rays"]
        if ref is None:
            ref = (accum.copy(), rays)
        assert rays == ref[1] and np.allclose(accum, ref[0], rtol=1e-5, atol=1e-5)
        scene.close()
    oi = rt.World("cornell_smoke").camera()
    ocam = S.make_camera(oi["lookfrom"], oi["lookat"], oi["field_of_view"], 1.0)
    r = rt.Renderer.new_with_rng(ocam, other, rt.BlackBackground(), rt.RenderingParams(4, 32, 32), rt.RecursiveRayTracer(50), rt.SeedableRngator(1))
    _, a = r.render_arrays()
    assert np.isfinite(a).all() and a.sum() > 0
    other.close()
