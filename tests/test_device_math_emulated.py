"""CPU-side checks of the DEVICE math: rt_device.cuh compiled for the host (tests/emul) against the oracle.

The container that runs `-m "not gpu"` has no GPU, so these tests exercise the very same functions the CUDA kernels
inline (hit_sphere, hit_box, closest_hit, medium_interval, texture_value, Philox, the samplers, integrate_item)
through a host build of the header.  The emulation is test infrastructure; the `-m gpu` tests repeat the checks
through the real C ABI on the B200."""
import ctypes as C

import numpy as np
import pytest

import mu_lambda_raytracer_b200 as rt
from mu_lambda_raytracer_b200 import abi
import parity_cases as PC
import support as S


@pytest.mark.parametrize("case", PC.SURFACE_CASES, ids=lambda c: c.__name__)
def test_primitive_hits_match_oracle(case):
    rng = np.random.default_rng(11)
    b, node, rays = case(rng, 100_000)
    desc = b.finish(node)
    es = S.EmulScene(desc, build_bvh=False)
    ow = S.OracleWorld(desc=desc)
    r = PC.compare_surface(es.intersect(rays, mode=1), ow.hit(rays, node=node), rays)
    assert r["hits"] > 0.1 * r["n"], r
    assert r["hard_mismatch"] == 0 and r["grazing_mismatch"] <= 20, r
    assert r["t_bad"] == 0 and r["n_bad"] == 0 and r["p_bad"] == 0 and r["ff_bad"] == 0 and r["uv_bad"] <= 2 and r["mat_bad"] == 0, r


@pytest.mark.parametrize("case", PC.MEDIUM_CASES, ids=lambda c: c.__name__)
def test_medium_interval_matches_oracle(case):
    rng = np.random.default_rng(12)
    b, node, rays = case(rng, 100_000)
    desc = b.finish(node)
    es = S.EmulScene(desc, build_bvh=False)
    ow = S.OracleWorld(desc=desc)
    g = es.intersect(rays, mode=2)
    ohit, ot = ow.medium_interval(rays, node)
    ghit = g["material"] >= 0
    dlen = np.linalg.norm(rays[:, 3:6], axis=1)
    scale = np.abs(rays[:, :3]).max(axis=1) + 1000.0
    both = (ohit == 1) & ghit
    assert both.sum() > (0.05 if "unordered" in case.__name__ else 0.2) * len(rays)  # degenerate sides: few two-crossing rays
    # disagreements only where the chord is within rounding of the 0.001 re-entry epsilon / zero length
    mism = (ohit == 1) != ghit
    assert mism.sum() <= 10
    assert np.all(np.abs(g["t"][both] - ot[both, 0]) * dlen[both] <= 2e-5 * scale[both])
    assert np.all(np.abs(g["u"][both] - ot[both, 1]) * dlen[both] <= 2e-5 * scale[both])


@pytest.mark.parametrize("name", ["simple", "random", "cornell_box", "final_scene"])
def test_bvh_closest_hit_matches_oracle_and_brute_force(name):
    import test_gpu_parity as G
    rng = np.random.default_rng(5)
    desc = rt.World(name).build(42)
    es = S.EmulScene(desc.ptr)
    ow = S.OracleWorld(name, 42)
    lo, hi, aim, spread = G.RAY_BOXES[name]
    rays = S.random_rays(60_000, rng, lo, hi, target=aim, spread=spread)
    g = es.intersect(rays, mode=0)
    G.check_hits(g, ow.hit(rays), rays)
    lin = es.intersect(rays, mode=1)  # brute force over the same primitives: identical up to ties
    assert np.array_equal(lin["material"] >= 0, g["material"] >= 0)
    hit = g["material"] >= 0
    assert np.allclose(lin["t"][hit], g["t"][hit], rtol=1e-6, atol=0)


def philox_reference(ctr, key):
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c, k = list(ctr), list(key)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
        k = [(k[0] + W0) & 0xFFFFFFFF, (k[1] + W1) & 0xFFFFFFFF]
    return c


def test_philox4x32_10_known_answers():
    # Random123's kat_vectors for philox4x32-10
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        assert tuple(philox_reference(ctr, key)) == want
        out = (C.c_uint32 * 4)()
        S.emul().emul_philox((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
        assert tuple(out) == want


def test_unit_ball_sampler_is_uniform():
    rng = np.random.default_rng(3)
    u = rng.random((200_000, 3)).astype(np.float32)
    p = np.zeros_like(u)
    S.emul().emul_unit_ball(u.ctypes.data, len(u), p.ctypes.data)
    r = np.linalg.norm(p, axis=1)
    assert r.max() <= 1.0 + 1e-6
    # uniform in the ball: r^3 uniform, direction isotropic
    assert abs(np.mean(r ** 3) - 0.5) < 0.005 and np.abs(p.mean(axis=0)).max() < 0.005
    assert np.abs(np.mean(p * p, axis=0) - 0.2).max() < 0.003  # E[x^2] = 1/5 for the unit ball


def test_camera_rays_follow_camera_rs():
    cam = S.make_camera((13, 2, 3), (0, 0, 0), 20.0, 1.5, aperture=0.1, focus_dist=10.0)
    p = abi.RtParams()
    p.width, p.height, p.samples_per_pixel, p.max_depth, p.seed = 1200, 800, 4, 50, 42
    n = 5000
    rng = np.random.default_rng(0)
    pixel = rng.integers(0, 1200 * 800, n).astype(np.int32)
    sample = rng.integers(0, 500, n).astype(np.int32)
    rays, us = np.zeros((n, 6), np.float32), np.zeros((n, 4), np.float32)
    S.emul().emul_generate_rays(C.byref(cam.c), C.byref(p), pixel.ctypes.data, sample.ctypes.data, n, rays.ctypes.data, us.ctypes.data)
    basis = (C.c_double * 19)()
    S.oracle().orc_camera_basis(C.byref(cam.c), basis)
    b = np.array(basis)
    origin, llc, hor, ver, cu, cv, lens = b[0:3], b[3:6], b[6:9], b[9:12], b[12:15], b[15:18], b[18]
    i, j = pixel % 1200, pixel // 1200
    s = (i + us[:, 0].astype(np.float64)) / 1199.0
    t = (j + us[:, 1].astype(np.float64)) / 799.0
    r = lens * np.sqrt(us[:, 2].astype(np.float64))
    phi = 2 * np.pi * us[:, 3].astype(np.float64)
    off = np.outer(r * np.cos(phi), cu) + np.outer(r * np.sin(phi), cv)
    want_o = origin + off
    want_d = llc + np.outer(s, hor) + np.outer(t, ver) - origin - off
    assert np.abs(rays[:, :3] - want_o).max() < 2e-6 * 13
    assert np.abs(rays[:, 3:] - want_d).max() < 2e-6 * 10
    assert us.min() >= 0.0 and us.max() < 1.0
    assert abs(us.mean() - 0.5) < 0.01
    # disk sample is uniform: E[r^2] = lens^2 / 2
    assert abs(np.mean(np.sum(off * off, axis=1)) / (lens * lens) - 0.5) < 0.02


def test_textures_match_oracle():
    fs = rt.World("final_scene").build(42)
    b = S.DescBuilder()
    odd, even = b.solid(0.2, 0.3, 0.1), b.solid(0.9, 0.9, 0.9)
    chk = b.checker(odd, even)
    noi = b.noise(0.1, perlin_from=fs.desc.perlins[0])
    noi4 = b.noise(4.0, perlin_from=fs.desc.perlins[0])
    img = b.image(S.earthmap())
    node = b.sphere((0, 0, 0), 1.0, b.lambertian(chk))
    desc = b.finish(node)
    es = S.EmulScene(desc, build_bvh=False)
    ow = S.OracleWorld(desc=desc)
    rng = np.random.default_rng(9)
    n = 20000
    uvp = np.zeros((n, 5))
    uvp[:, 0:2] = rng.uniform(-0.1, 1.1, (n, 2))
    uvp[:, 2:5] = rng.uniform(-400, 600, (n, 3))
    uvp = uvp.astype(np.float32).astype(np.float64)
    for tex, tol, frac in ((chk, 1e-6, 0.002), (noi, 2e-3, 0.0), (img, 1e-6, 0.001)):
        want = np.zeros((n, 3))
        assert S.oracle().orc_texture_value(ow.h, tex, uvp.ctypes.data, n, want.ctypes.data) == 0
        got = np.zeros((n, 3), np.float32)
        u32 = uvp.astype(np.float32)
        S.emul().emul_texture_value_batch(es.h, tex, u32.ctypes.data, n, got.ctypes.data)
        bad = np.abs(got - want).max(axis=1) > tol
        assert bad.mean() <= frac, (tex, bad.mean(), np.abs(got - want).max())
    # scale 4 noise on small coordinates (two_spheres / simple_light)
    uvp[:, 2:5] = rng.uniform(-12, 12, (n, 3))
    uvp = uvp.astype(np.float32).astype(np.float64)
    want = np.zeros((n, 3))
    S.oracle().orc_texture_value(ow.h, noi4, uvp.ctypes.data, n, want.ctypes.data)
    got = np.zeros((n, 3), np.float32)
    u32 = uvp.astype(np.float32)
    S.emul().emul_texture_value_batch(es.h, noi4, u32.ctypes.data, n, got.ctypes.data)
    assert np.abs(got - want).max() < 2e-3


@pytest.mark.parametrize("name,aspect,kw", [("random", 1.5, dict(aperture=0.1, focus_dist=10.0)), ("cornell_smoke", 1.0, {}),
                                            ("final_scene", 1.0, {}), ("simple", 16 / 9, {}), ("cornell_box", 1.0, {}),
                                            ("simple_light", 1.5, {}), ("random_chk", 1.5, {}), ("earth", 1.5, {})])
def test_emulated_render_sits_at_oracle_noise_floor(name, aspect, kw):
    world = rt.World(name)
    desc = world.build(42)
    es = S.EmulScene(desc.ptr)
    ow = S.OracleWorld(name, 42)
    cam = S.make_camera(ow.lookfrom, ow.lookat, ow.vfov, aspect, **kw)
    W, spp = 64, 48
    H = int(W / aspect)
    a1, _, c1, _ = ow.render(cam.c, W, H, spp, render_seed=42)
    a2, _, _, _ = ow.render(cam.c, W, H, spp, render_seed=1042)
    e, rays = es.render(cam, W, H, spp, seed=42)
    disp = lambda a: np.sqrt(np.clip(a / spp, 0.0, 1.0))
    rm = lambda x, y: float(np.sqrt(np.mean((x - y) ** 2)))
    floor = rm(disp(a1), disp(a2))
    got = 0.5 * (rm(disp(e.astype(np.float64)), disp(a1)) + rm(disp(e.astype(np.float64)), disp(a2)))
    assert got <= 1.1 * floor, (got, floor)
    assert abs(rays / (W * H * spp) - c1[1] / c1[0]) < 0.04 * c1[1] / c1[0]
