"""GPU parity, level 2 (BASELINE.json north_star): every primitive kind, the medium intervals, the textures and the
camera — the CUDA kernels called through the C ABI against the oracle on 1 M random rays with NON-unit directions
(hit/miss identical outside a grazing band, t and normal within 1e-5 relative), plus the reference's own unit-test
known answers (src/shapes.rs:204-213, src/bhv.rs:173-220) replayed on the device."""
import ctypes as C

import numpy as np
import pytest

import mu_lambda_raytracer_b200 as rt
from mu_lambda_raytracer_b200 import abi
import parity_cases as PC
import support as S

pytestmark = pytest.mark.gpu
N_RAYS = 1_000_000


def _scene(desc_ptr):
    d = rt.SceneDescription(desc_ptr, owned=False)
    return rt.Scene(d)


def _intersect(scene, rays, node):
    r32 = np.ascontiguousarray(rays, dtype=np.float32)
    hits = (abi.RtHit * len(rays))()
    abi.check(abi.load().rt_intersect_batch(scene.handle, node, r32.ctypes.data, len(rays), hits))
    return np.ctypeslib.as_array(hits).copy()


@pytest.mark.parametrize("case", PC.SURFACE_CASES, ids=lambda c: c.__name__)
def test_primitive_hits_match_oracle_1m_rays(case):
    rng = np.random.default_rng(21)
    b, node, rays = case(rng, N_RAYS)
    desc = b.finish(node)
    scene = _scene(desc)
    ow = S.OracleWorld(desc=desc)
    g = _intersect(scene, rays, node)
    o = ow.hit(rays, node=node)
    r = PC.compare_surface(g, o, rays)
    assert r["hits"] > 0.1 * r["n"], r
    if r["hard_mismatch"]:  # leave the evidence in the log
        for i in np.where((o["hit"] == 1) != (g["material"] >= 0))[0][:8]:
            print("mismatch ray", i, rays[i].tolist(), "oracle", o[i], "device", g[i])
    assert r["hard_mismatch"] == 0 and r["grazing_mismatch"] <= 200, r
    assert r["t_bad"] == 0 and r["n_bad"] == 0 and r["p_bad"] == 0 and r["ff_bad"] == 0 and r["mat_bad"] == 0, r
    assert r["uv_bad"] <= 20, r
    scene.close()


@pytest.mark.parametrize("case", PC.MEDIUM_CASES, ids=lambda c: c.__name__)
def test_medium_interval_matches_oracle_1m_rays(case):
    rng = np.random.default_rng(22)
    b, node, rays = case(rng, N_RAYS)
    desc = b.finish(node)
    scene = _scene(desc)
    ow = S.OracleWorld(desc=desc)
    g = _intersect(scene, rays, node)
    ohit, ot = ow.medium_interval(rays, node)
    ghit = g["material"] >= 0
    dlen = np.linalg.norm(rays[:, 3:6], axis=1)
    scale = np.abs(rays[:, :3]).max(axis=1) + 1000.0
    both = (ohit == 1) & ghit
    assert both.sum() > (0.05 if "unordered" in case.__name__ else 0.2) * len(rays)  # degenerate sides: few two-crossing rays
    assert ((ohit == 1) != ghit).sum() <= 100  # only chords within rounding of the 0.001 re-entry epsilon
    assert np.all(np.abs(g["t"][both] - ot[both, 0]) * dlen[both] <= 2e-5 * scale[both])
    assert np.all(np.abs(g["u"][both] - ot[both, 1]) * dlen[both] <= 2e-5 * scale[both])
    scene.close()


def test_reference_sphere_uv_known_answers_on_device():
    """src/shapes.rs:204-213 (test_sphere_uv): the six axis points of the unit sphere, here as hits of rays fired
    at them from outside; the reference asserts f64 equality, the device is f32 (atan2f/acosf): 1e-6."""
    cases = [((1, 0, 0), (0.5, 0.5)), ((0, 1, 0), (0.5, 1.0)), ((0, 0, 1), (0.25, 0.5)), ((-1, 0, 0), (0.0, 0.5)),
             ((0, -1, 0), (0.5, 0.0)), ((0, 0, -1), (0.75, 0.5))]
    b = S.DescBuilder()
    node = b.sphere((0.0, 0.0, 0.0), 1.0, b.lambertian(b.solid(0.5, 0.5, 0.5)))
    scene = _scene(b.finish(node))
    rays = np.array([[3 * n[0], 3 * n[1], 3 * n[2], -n[0], -n[1], -n[2], 0.001, np.inf] for n, _ in cases], dtype=np.float64)
    g = _intersect(scene, rays, node)
    for k, (n, (u, v)) in enumerate(cases):
        assert g["material"][k] >= 0 and abs(g["t"][k] - 2.0) < 1e-6
        du = abs(g["u"][k] - u)
        assert min(du, 1.0 - du) < 1e-6 and abs(g["v"][k] - v) < 1e-6, (n, g["u"][k], g["v"][k])
        assert np.allclose(g["normal"][k], n, atol=1e-6) and g["front_face"][k] == 1
    scene.close()


def test_reference_aabb_known_answers_on_device():
    """src/bhv.rs:173-220 (test_edge_parallel*, test_face_parallel*): the same rays against the same box, given to the
    device as a Block (its six rects answer exactly like the box test for these rays), and through a one-leaf BVH."""
    b = S.DescBuilder()
    blk = b.block((1.0, 1.0, 1.0), (2.0, 2.0, 2.0), b.lambertian(b.solid(0.5, 0.5, 0.5)))
    root = b.group(abi.RT_NODE_BVH, [blk])
    scene = _scene(b.finish(root))
    inf = np.inf
    rays = np.array([[0, 0, 0, 1, 1, 1, 0.0, inf], [1.0001, 0, 1.0001, 0, 1, 0, 0.0, inf], [0.99999, 0, 0.9999, 0, 1, 0, 0.0, inf],
                     [1.5, 0, 1.0001, 0, 3, 0, 0.0, inf], [1.5, 0, 0.9999, 0, 3, 0, 0.0, inf]], dtype=np.float64)
    want = [True, True, False, True, False]
    for node in (blk, -1):  # brute force over the sub-tree, then the BVH of the whole description
        g = _intersect(scene, rays, node)
        assert list(g["material"] >= 0) == want, node
    g = _intersect(scene, rays, -1)
    assert abs(g["t"][0] - 1.0) < 1e-6 and abs(g["t"][3] - 1.0 / 3.0) < 1e-6
    scene.close()


def test_textures_match_oracle_on_device():
    fs = rt.World("final_scene").build(42)
    b = S.DescBuilder()
    odd, even = b.solid(0.2, 0.3, 0.1), b.solid(0.9, 0.9, 0.9)
    chk = b.checker(odd, even)
    noi = b.noise(0.1, perlin_from=fs.desc.perlins[0])
    noi4 = b.noise(4.0, perlin_from=fs.desc.perlins[0])
    img = b.image(S.earthmap())
    node = b.sphere((0, 0, 0), 1.0, b.lambertian(chk))
    desc = b.finish(node)
    scene = _scene(desc)
    ow = S.OracleWorld(desc=desc)
    rng = np.random.default_rng(9)
    n = 200_000
    uvp = np.zeros((n, 5))
    uvp[:, 0:2] = rng.uniform(-0.1, 1.1, (n, 2))
    uvp[:, 2:5] = rng.uniform(-400, 600, (n, 3))
    uvp = uvp.astype(np.float32).astype(np.float64)

    def both(tex, pts):
        want = np.zeros((n, 3))
        assert S.oracle().orc_texture_value(ow.h, tex, pts.ctypes.data, n, want.ctypes.data) == 0
        got = np.zeros((n, 3), np.float32)
        u32 = pts.astype(np.float32)
        abi.check(abi.load().rt_texture_value_batch(scene.handle, tex, u32.ctypes.data, n, got.ctypes.data))
        return got, want

    for tex, tol, frac in ((chk, 1e-6, 0.002), (noi, 3e-3, 0.0), (img, 1e-6, 0.001)):
        got, want = both(tex, uvp)
        bad = np.abs(got - want).max(axis=1) > tol
        assert bad.mean() <= frac, (tex, bad.mean(), np.abs(got - want).max())
    uvp[:, 2:5] = rng.uniform(-12, 12, (n, 3))
    uvp = uvp.astype(np.float32).astype(np.float64)
    got, want = both(noi4, uvp)
    assert np.abs(got - want).max() < 3e-3
    scene.close()


def test_camera_rays_follow_camera_rs_on_device():
    cam = S.make_camera((13, 2, 3), (0, 0, 0), 20.0, 1.5, aperture=0.1, focus_dist=10.0)
    p = abi.RtParams()
    p.width, p.height, p.samples_per_pixel, p.max_depth, p.seed = 1200, 800, 4, 50, 42
    n = 100_000
    rng = np.random.default_rng(0)
    pixel = rng.integers(0, 1200 * 800, n).astype(np.int32)
    sample = rng.integers(0, 500, n).astype(np.int32)
    rays, us = np.zeros((n, 6), np.float32), np.zeros((n, 4), np.float32)
    abi.check(abi.load().rt_generate_rays(C.byref(cam.c), C.byref(p), pixel.ctypes.data, sample.ctypes.data, n, rays.ctypes.data, us.ctypes.data))
    # the same numbers as the host build of the device header (Philox counters are (pixel, sample, 0))
    rays_e, us_e = np.zeros((n, 6), np.float32), np.zeros((n, 4), np.float32)
    S.emul().emul_generate_rays(C.byref(cam.c), C.byref(p), pixel.ctypes.data, sample.ctypes.data, n, rays_e.ctypes.data, us_e.ctypes.data)
    assert np.array_equal(us, us_e)
    assert np.abs(rays - rays_e).max() < 1e-5
    basis = (C.c_double * 19)()
    S.oracle().orc_camera_basis(C.byref(cam.c), basis)
    bb = np.array(basis)
    origin, llc, hor, ver, cu, cv, lens = bb[0:3], bb[3:6], bb[6:9], bb[9:12], bb[12:15], bb[15:18], bb[18]
    i, j = pixel % 1200, pixel // 1200
    s = (i + us[:, 0].astype(np.float64)) / 1199.0
    t = (j + us[:, 1].astype(np.float64)) / 799.0
    r = lens * np.sqrt(us[:, 2].astype(np.float64))
    phi = 2 * np.pi * us[:, 3].astype(np.float64)
    off = np.outer(r * np.cos(phi), cu) + np.outer(r * np.sin(phi), cv)
    assert np.abs(rays[:, :3] - (origin + off)).max() < 2e-6 * 13
    assert np.abs(rays[:, 3:] - (llc + np.outer(s, hor) + np.outer(t, ver) - origin - off)).max() < 2e-6 * 10
    assert us.min() >= 0.0 and us.max() < 1.0 and abs(us.mean() - 0.5) < 0.005
