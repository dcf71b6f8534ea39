"""Material::scatter / emit parity (materials.rs:25-127, volumes.rs:77-83): the device's `scatter` on caller-supplied hits
against the oracle's.  Where the reference's sampler is a function of ONE uniform (Dielectric) or of none (Metal with
fuzz 0, DiffuseLight, absorbed Metal) the two are compared event by event, the device being fed the very uniform the
oracle's PCG stream yields.  Where the reference rejection-samples (Lambertian, fuzzy Metal, Isotropic: vec.rs:23-43)
and the device samples the same distribution directly, the two sample sets are compared through their moments.
Shared by the host-emulation test (CPU suite) and the C-ABI test on the GPU."""
import ctypes as C

import numpy as np

from mu_lambda_raytracer_b200 import abi
import support as S


def build_scene():
    b = S.DescBuilder()
    mats = {
        "lambertian": b.lambertian(b.solid(0.7, 0.4, 0.2)),
        "lambertian_checker": b.lambertian(b.checker(b.solid(0.2, 0.3, 0.1), b.solid(0.9, 0.9, 0.9))),
        "metal_mirror": b.material(abi.RT_MAT_METAL, albedo=(0.8, 0.6, 0.2), fuzz=0.0),
        "metal_fuzzy": b.material(abi.RT_MAT_METAL, albedo=(0.7, 0.7, 0.9), fuzz=0.4),
        "glass": b.material(abi.RT_MAT_DIELECTRIC, ior=1.5),
        "light": b.material(abi.RT_MAT_DIFFUSE_LIGHT, b.solid(7.0, 6.0, 5.0)),
        "isotropic": b.material(abi.RT_MAT_ISOTROPIC, b.solid(0.2, 0.4, 0.9)),
    }
    node = b.sphere((0, 0, 0), 1.0, mats["lambertian"])
    return b, b.finish(node), mats


def random_events(n, rng, grazing_ok=True):
    """hits on a unit-ish sphere seen by rays with NON-unit directions: (ray o, d, p, face-forwarded normal, u, v, front)"""
    nrm = rng.normal(size=(n, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    front = rng.integers(0, 2, n).astype(np.float64)
    # the stored normal always opposes the ray (hittable.rs:18-30); front_face only records which side that was
    flip = np.sum(d * nrm, axis=1) > 0
    nrm[flip] *= -1
    d *= rng.uniform(0.05, 3.0, size=(n, 1))
    p = rng.uniform(-5, 5, size=(n, 3))
    o = p - d * rng.uniform(0.5, 4.0, size=(n, 1))
    ev = np.zeros((n, 15))
    ev[:, 0:3], ev[:, 3:6], ev[:, 6:9], ev[:, 9:12] = o, d, p, nrm
    ev[:, 12:14] = rng.uniform(0, 1, (n, 2))
    ev[:, 14] = front
    return ev.astype(np.float32).astype(np.float64)


def oracle_scatter(ow, material, ev, seed_base=1000):
    out = np.zeros((len(ev), 11))
    ev = np.ascontiguousarray(ev)
    assert S.oracle().orc_scatter_batch(ow.h, material, ev.ctypes.data, len(ev), seed_base, out.ctypes.data) == 0
    return out


def device_inputs(ev, material, uniforms):
    arr = (abi.RtScatterIn * len(ev))()
    a = np.ctypeslib.as_array(arr)
    a["ray_origin"], a["ray_dir"], a["p"], a["normal"] = ev[:, 0:3], ev[:, 3:6], ev[:, 6:9], ev[:, 9:12]
    a["u"], a["v"], a["front_face"], a["material"] = ev[:, 12], ev[:, 13], ev[:, 14].astype(np.int32), material
    a["uniform"] = uniforms
    return arr


def check_all(run_device, n=200_000):
    """run_device(desc, RtScatterIn array) -> structured RtScatterOut array"""
    b, desc, mats = build_scene()
    ow = S.OracleWorld(desc=desc)
    rng = np.random.default_rng(77)
    ev = random_events(n, rng)
    unit = lambda v: v / np.linalg.norm(v, axis=1, keepdims=True)
    ud, nrm = unit(ev[:, 3:6]), ev[:, 9:12]
    refl = ud - 2 * np.sum(ud * nrm, axis=1, keepdims=True) * nrm

    def dev(material, uniforms):
        return run_device(desc, device_inputs(ev, material, uniforms))

    U = rng.uniform(0, 1, (n, 4)).astype(np.float32)

    # ---- Metal, fuzz 0 (materials.rs:51-61): the mirror direction, event by event
    o, g = oracle_scatter(ow, mats["metal_mirror"], ev), dev(mats["metal_mirror"], U)
    assert np.array_equal(g["scattered"], o[:, 0].astype(np.int32)) and g["scattered"].all()  # the stored normal opposes the ray
    assert np.abs(g["dir"] - o[:, 4:7]).max() < 2e-6 and np.abs(g["attenuation"] - o[:, 1:4]).max() < 1e-6

    # ---- DiffuseLight (materials.rs:119-127): never scatters, emits from BOTH faces
    o, g = oracle_scatter(ow, mats["light"], ev), dev(mats["light"], U)
    assert not g["scattered"].any() and not o[:, 0].any()
    assert np.array_equal(g["emitted"], np.tile(np.float32([7, 6, 5]), (n, 1))) and np.abs(g["emitted"] - o[:, 7:10]).max() == 0
    assert (ev[:, 14] == 0).sum() > n // 3  # back faces were in the batch

    # ---- Dielectric (materials.rs:88-106): the device gets the uniform the oracle's stream yields -> same branch, same ray
    o = oracle_scatter(ow, mats["glass"], ev)
    Ud = U.copy()
    Ud[:, 3] = o[:, 10].astype(np.float32)
    g = dev(mats["glass"], Ud)
    assert g["scattered"].all() and o[:, 0].all() and np.array_equal(g["attenuation"], np.ones((n, 3), np.float32))
    cos_t = np.minimum(-np.sum(ud * nrm, axis=1), 1.0)
    ratio = np.where(ev[:, 14] != 0, 1 / 1.5, 1.5)
    r0 = ((1 - ratio) / (1 + ratio)) ** 2
    refl_prob = r0 + (1 - r0) * (1 - cos_t) ** 5
    sin_t = np.sqrt(1 - cos_t ** 2)
    # events whose decision sits within f32 rounding of the threshold may legitimately go the other way
    near = (np.abs(refl_prob - o[:, 10]) < 2e-6) | (np.abs(ratio * sin_t - 1.0) < 2e-6)
    err = np.abs(g["dir"] - o[:, 4:7]).max(axis=1)
    # 1e-5 (north_star's bar for f32 against f64), plus the conditioning of refract()'s sqrt(|1 - |r_perp|^2|)
    # (materials.rs:66) next to the critical angle, where a rounding of 1e-7 in r_perp is amplified by 1 / (2 sqrt(k))
    k = np.abs(1.0 - (ratio * sin_t) ** 2)
    tol = 1e-5 + 1e-6 / np.sqrt(np.maximum(k, 1e-12))
    assert ((err > tol) & ~near).sum() == 0, (err[~near].max(), ((err > tol) & ~near).sum())
    is_refl = np.abs(o[:, 4:7] - refl).max(axis=1) < 1e-12
    assert 0.05 < is_refl.mean() < 0.95  # both branches (and total internal reflection) were exercised
    assert (ratio * sin_t > 1.0).sum() > n // 50

    # ---- Metal, fuzz 0.4: reflect + fuzz * in-ball; absorbed (emit = 0) when it ends below the surface
    o, g = oracle_scatter(ow, mats["metal_fuzzy"], ev), dev(mats["metal_fuzzy"], U)
    for name, alive, dirs in (("oracle", o[:, 0] != 0, o[:, 4:7]), ("device", g["scattered"] != 0, g["dir"].astype(np.float64))):
        ball = (dirs[alive] - refl[alive]) / 0.4
        assert np.linalg.norm(ball, axis=1).max() < 1.0 + 1e-5, name
        assert np.all(np.sum(dirs[alive] * nrm[alive], axis=1) > 0), name
    assert abs((g["scattered"] != 0).mean() - (o[:, 0] != 0).mean()) < 4 * np.sqrt(0.25 / n) * 2
    assert 0.02 < (g["scattered"] == 0).mean() < 0.5 and not g["emitted"][g["scattered"] == 0].any()
    # the absorbed set depends on the sample, so compare the samples through a reflection-frame statistic both keep: the
    # component of (dir - refl) along the normal, over ALL events where the mirror ray leaves at > 0.4 (never absorbed)
    safe = np.sum(refl * nrm, axis=1) > 0.4
    for k in range(3):
        mo = ((o[safe, 4 + k] - refl[safe, k]) / 0.4)
        mg = ((g["dir"][safe, k] - refl[safe, k]) / 0.4)
        se = np.sqrt(0.2 / safe.sum())
        assert abs(mo.mean()) < 5 * se and abs(mg.mean()) < 5 * se
        assert abs((mo ** 2).mean() - 0.2) < 5 * se and abs((mg ** 2).mean() - 0.2) < 5 * se  # E[x^2] of the uniform ball = 1/5

    # ---- Lambertian (materials.rs:25-34): normal + in-ball point flipped into the normal's hemisphere (vec.rs:36-43)
    for key in ("lambertian", "lambertian_checker"):
        o, g = oracle_scatter(ow, mats[key], ev), dev(mats[key], U)
        assert g["scattered"].all() and o[:, 0].all()
        if key == "lambertian":
            assert np.abs(g["attenuation"] - np.float32([0.7, 0.4, 0.2])).max() < 1e-7
        else:  # Checker on the hit point (textures.rs:40-49): same side except within rounding of a zero of the sines
            assert (np.abs(g["attenuation"] - o[:, 1:4]).max(axis=1) > 1e-6).mean() < 0.002
        stats = {}
        for name, dirs in (("oracle", o[:, 4:7]), ("device", g["dir"].astype(np.float64))):
            bvec = dirs - nrm
            along = np.sum(bvec * nrm, axis=1)
            r = np.linalg.norm(bvec, axis=1)
            assert r.max() < 1.0 + 1e-5 and along.min() > -1e-6, name
            tang = bvec - along[:, None] * nrm
            stats[name] = (along.mean(), r.mean(), (along ** 2).mean(), (r ** 2).mean(), np.abs(tang.mean(axis=0)).max())
        # expected moments of the flipped uniform ball, each within 5 standard errors (variances: upper bounds)
        for k, (expect, var) in enumerate(((3 / 8, 0.06), (3 / 4, 0.04), (1 / 5, 0.05), (3 / 5, 0.07), (0.0, 0.2))):
            tol = 5 * np.sqrt(var / n)
            assert abs(stats["oracle"][k] - expect) < tol and abs(stats["device"][k] - expect) < tol, (key, k, stats)

    # ---- Isotropic (volumes.rs:77-83): the raw in-ball point, NOT normalised
    o, g = oracle_scatter(ow, mats["isotropic"], ev), dev(mats["isotropic"], U)
    assert g["scattered"].all() and np.abs(g["attenuation"] - np.float32([0.2, 0.4, 0.9])).max() < 1e-7
    for name, dirs in (("oracle", o[:, 4:7]), ("device", g["dir"].astype(np.float64))):
        r = np.linalg.norm(dirs, axis=1)
        se = 5 / np.sqrt(n)
        assert r.max() < 1.0 + 1e-6 and abs(r.mean() - 0.75) < 0.2 * se and abs((r ** 3 < 0.5).mean() - 0.5) < 0.5 * se, name  # radius^3 uniform
        assert np.abs(dirs.mean(axis=0)).max() < 0.45 * se and np.abs((dirs ** 2).mean(axis=0) - 0.2).max() < 0.22 * se, name
    return True
