"""Per-primitive parity cases shared by the CPU emulation tests and the GPU tests: each case builds a small
description, a batch of rays with NON-unit directions (origins inside and outside, sizes from r = 0.2 to 1000),
and names the node to query."""
import numpy as np

from mu_lambda_raytracer_b200 import abi
import support as S


def _mat(b):
    return b.lambertian(b.solid(0.5, 0.5, 0.5))


def case_small_sphere(rng, n):
    b = S.DescBuilder()
    node = b.sphere((3.0, 0.2, -2.0), 0.2, _mat(b))
    rays = S.random_rays(n, rng, [-5, -3, -8], [9, 4, 4], target=[3.0, 0.2, -2.0], spread=[0.3, 0.3, 0.3])
    return b, node, rays


def case_sphere_from_inside(rng, n):
    b = S.DescBuilder()
    node = b.sphere((260.0, 150.0, 45.0), 50.0, _mat(b))
    rays = S.random_rays(n, rng, [230, 120, 15], [290, 180, 75])
    return b, node, rays


def case_negative_radius_sphere(rng, n):
    b = S.DescBuilder()
    node = b.sphere((-1.0, 0.0, -1.0), -0.4, b.material(abi.RT_MAT_DIELECTRIC, ior=1.5))
    rays = S.random_rays(n, rng, [-3, -2, -3], [1, 2, 1], target=[-1, 0, -1], spread=[0.5, 0.5, 0.5])
    return b, node, rays


def case_ground_sphere(rng, n):  # r = 1000 seen from just above its surface (the `random` world)
    b = S.DescBuilder()
    node = b.sphere((0.0, -1000.0, 0.0), 1000.0, _mat(b))
    rays = S.random_rays(n, rng, [-15, 0.01, -15], [15, 4, 15], target=[0, -0.5, 0], spread=[14, 1.5, 14])
    return b, node, rays


def case_fog_sphere_from_inside(rng, n):  # r = 1000 from deep inside (final_scene's fog boundary)
    b = S.DescBuilder()
    node = b.sphere((0.0, 0.0, 0.0), 1000.0, _mat(b))
    rays = S.random_rays(n, rng, [-500, 0, -600], [500, 550, 500])
    return b, node, rays


def case_xy_rect(rng, n):
    b = S.DescBuilder()
    node = b.rect(abi.RT_NODE_XYRECT, 3.0, 5.0, 1.0, 3.0, -2.0, _mat(b))
    rays = S.random_rays(n, rng, [-2, -3, -9], [9, 7, 5], target=[4, 2, -2], spread=[1.6, 1.6, 0.0])
    return b, node, rays


def case_xz_rect(rng, n):
    b = S.DescBuilder()
    node = b.rect(abi.RT_NODE_XZRECT, 123.0, 423.0, 147.0, 412.0, 554.0, _mat(b))
    rays = S.random_rays(n, rng, [0, 0, 0], [555, 900, 555], target=[273, 554, 280], spread=[200, 0, 180])
    return b, node, rays


def case_yz_rect(rng, n):
    b = S.DescBuilder()
    node = b.rect(abi.RT_NODE_YZRECT, 0.0, 555.0, 0.0, 555.0, 555.0, _mat(b))
    rays = S.random_rays(n, rng, [0, 0, 0], [900, 555, 555], target=[555, 278, 278], spread=[0, 330, 330])
    return b, node, rays


def case_block(rng, n):
    b = S.DescBuilder()
    node = b.block((-300.0, 0.0, 100.0), (-200.0, 53.5, 200.0), _mat(b))
    rays = S.random_rays(n, rng, [-500, -50, -100], [0, 300, 400], target=[-250, 27, 150], spread=[60, 35, 60])
    return b, node, rays


def case_block_from_inside(rng, n):
    b = S.DescBuilder()
    node = b.block((0.0, 0.0, 0.0), (165.0, 330.0, 165.0), _mat(b))
    rays = S.random_rays(n, rng, [1, 1, 1], [164, 329, 164])
    return b, node, rays


def case_rotated_translated_block(rng, n):  # cornell_box's tall block
    b = S.DescBuilder()
    blk = b.block((0.0, 0.0, 0.0), (165.0, 330.0, 165.0), _mat(b))
    node = b.translate((265.0, 0.0, 295.0), b.rotate(1, 15.0, blk))
    rays = S.random_rays(n, rng, [0, 0, -800], [555, 555, 555], target=[340, 165, 380], spread=[120, 180, 120])
    return b, node, rays


def case_rotated_only_block(rng, n):  # Rotate outermost: the reference's face-forward quirk is observable
    b = S.DescBuilder()
    blk = b.block((-40.0, -20.0, -30.0), (50.0, 60.0, 70.0), b.material(abi.RT_MAT_DIELECTRIC, ior=1.5))
    node = b.rotate(1, -18.0, blk)
    rays = S.random_rays(n, rng, [-200, -100, -200], [200, 150, 200], target=[0, 20, 20], spread=[60, 50, 60])
    return b, node, rays


def case_rotate_x_translate_sphere(rng, n):  # a sphere under a rotation about X and a translation: uv must follow
    b = S.DescBuilder()
    sp = b.sphere((10.0, 5.0, -3.0), 4.0, _mat(b))
    node = b.translate((-2.0, 7.0, 1.0), b.rotate(0, 33.0, sp))
    rays = S.random_rays(n, rng, [-40, -40, -40], [40, 40, 40], target=[8, 13, 1], spread=[5, 5, 5])
    return b, node, rays


def case_sphere_medium(rng, n):  # final_scene's smoke ball
    b = S.DescBuilder()
    sp = b.sphere((360.0, 150.0, 145.0), 70.0, b.material(abi.RT_MAT_DIELECTRIC, ior=1.5))
    node = b.medium(sp, 0.2, (0.2, 0.4, 0.9))
    rays = S.random_rays(n, rng, [200, 0, 0], [520, 300, 300], target=[360, 150, 145], spread=[80, 80, 80])
    return b, node, rays


def case_fog_medium(rng, n):  # final_scene's global fog: camera inside an r = 1000 boundary
    b = S.DescBuilder()
    sp = b.sphere((0.0, 0.0, 0.0), 1000.0, b.material(abi.RT_MAT_DIELECTRIC, ior=1.5))
    node = b.medium(sp, 0.0001, (1, 1, 1))
    rays = S.random_rays(n, rng, [-500, 0, -600], [500, 550, 500], tmax=700.0)
    return b, node, rays


def case_box_medium(rng, n):  # cornell_smoke's rotated box of smoke
    b = S.DescBuilder()
    blk = b.block((0.0, 0.0, 0.0), (165.0, 165.0, 165.0), _mat(b))
    node = b.medium(b.translate((130.0, 0.0, 65.0), b.rotate(1, -18.0, blk)), 0.01, (1, 1, 1))
    rays = S.random_rays(n, rng, [0, 0, -800], [555, 555, 555], target=[210, 82, 150], spread=[110, 100, 110])
    return b, node, rays


def case_unordered_block(rng, n):
    """Block::new with corners that are not ordered (shapes.rs:173-186 + AARect::new's one-sided normalisation,
    aarects.rs:31-43): six rects, some of them degenerate — whatever the reference hits, the device must hit"""
    b = S.DescBuilder()
    node = b.block((5.0, 3.0, 4.0), (1.0, 6.0, 2.0), _mat(b))
    rays = S.random_rays(n, rng, [-6, -4, -5], [12, 13, 11], target=[3, 4.5, 3], spread=[2.6, 2.2, 1.6])
    return b, node, rays


def case_two_sphere_medium(rng, n):
    """ConstantMedium<O: Hittable> (volumes.rs:7-11) with a HittableList of two overlapping spheres as boundary: the
    medium occupies h1..h2 = the first two crossings of the LIST (volumes.rs:27-34)"""
    b = S.DescBuilder()
    m = b.material(abi.RT_MAT_DIELECTRIC, ior=1.5)
    boundary = b.group(abi.RT_NODE_LIST, [b.sphere((0.0, 0.0, 0.0), 30.0, m), b.sphere((35.0, 5.0, 0.0), 25.0, m)])
    node = b.medium(boundary, 0.05, (0.8, 0.8, 0.8))
    rays = S.random_rays(n, rng, [-90, -70, -70], [110, 70, 70], target=[15, 2, 0], spread=[45, 30, 30])
    return b, node, rays


def case_sphere_and_box_medium(rng, n):  # a BVH of a sphere and a rotated, translated block as boundary
    b = S.DescBuilder()
    m = b.material(abi.RT_MAT_DIELECTRIC, ior=1.5)
    blk = b.translate((20.0, -10.0, 5.0), b.rotate(1, 25.0, b.block((0.0, 0.0, 0.0), (40.0, 30.0, 20.0), m)))
    boundary = b.group(abi.RT_NODE_BVH, [b.sphere((0.0, 0.0, 0.0), 22.0, m), blk])
    node = b.medium(boundary, 0.02, (1, 1, 1))
    rays = S.random_rays(n, rng, [-80, -60, -60], [110, 70, 80], target=[20, 5, 10], spread=[40, 25, 25])
    return b, node, rays


def case_unordered_block_medium(rng, n):  # the six-rect form of a block as a medium boundary
    b = S.DescBuilder()
    node = b.medium(b.block((50.0, 0.0, 40.0), (10.0, 60.0, 5.0), _mat(b)), 0.03, (1, 1, 1))
    rays = S.random_rays(n, rng, [-60, -50, -60], [120, 110, 100], target=[30, 30, 22], spread=[26, 35, 22])
    return b, node, rays


SURFACE_CASES = [case_small_sphere, case_sphere_from_inside, case_negative_radius_sphere, case_ground_sphere,
                 case_fog_sphere_from_inside, case_xy_rect, case_xz_rect, case_yz_rect, case_block, case_block_from_inside,
                 case_rotated_translated_block, case_rotated_only_block, case_rotate_x_translate_sphere, case_unordered_block]
MEDIUM_CASES = [case_sphere_medium, case_fog_medium, case_box_medium, case_two_sphere_medium, case_sphere_and_box_medium,
                case_unordered_block_medium]


def compare_surface(g, o, rays, grazing=0.02, tol=1e-5):
    """GPU/emulated RtHit array `g` vs oracle OrcHit array `o`.  Returns a dict of violation counts."""
    ohit, ghit = o["hit"] == 1, g["material"] >= 0
    dlen = np.linalg.norm(rays[:, 3:6], axis=1)
    dirn = rays[:, 3:6] / dlen[:, None]
    cos = np.abs(np.sum(dirn * o["normal"], axis=1))
    cos_g = np.abs(np.sum(dirn * g["normal"].astype(np.float64), axis=1))  # where only the device hit, judge grazing by ITS normal
    mism = ohit != ghit
    # ... or passes within f32 rounding of the primitive's rim (u or v at 0 or 1: the edge of a rect / box face)
    rim = lambda h: (np.minimum(h["u"], 1.0 - h["u"]) < 2e-5) | (np.minimum(h["v"], 1.0 - h["v"]) < 2e-5)
    hard = mism & ~(ohit & ((cos < grazing) | rim(o))) & ~(ghit & ~ohit & ((cos_g < grazing) | rim(g)))
    both = ohit & ghit & (cos >= grazing)
    scale = np.abs(rays[:, :3]).max(axis=1) + np.abs(o["t"]) * dlen + 1.0
    err = np.abs(g["t"].astype(np.float64) - o["t"]) * dlen / scale
    nerr = np.abs(g["normal"].astype(np.float64) - o["normal"]).max(axis=1)
    perr = np.abs(g["p"].astype(np.float64) - o["p"]).max(axis=1) / scale
    uverr = np.maximum(np.abs(g["u"] - o["u"]), np.abs(g["v"] - o["v"]))
    uverr = np.minimum(uverr, np.abs(1.0 - uverr))  # u wraps at the sphere's seam
    return {
        "n": len(rays), "hits": int(ohit.sum()), "hard_mismatch": int(hard.sum()), "grazing_mismatch": int((mism & ~hard).sum()),
        "t_bad": int((both & (err > tol)).sum()), "t_max": float(err[both].max()) if both.any() else 0.0,
        "n_bad": int((both & (nerr > 1e-4)).sum()), "p_bad": int((both & (perr > 2 * tol)).sum()),
        "ff_bad": int((both & (g["front_face"] != o["front_face"])).sum()),
        "uv_bad": int((both & (uverr > 2e-4)).sum()), "mat_bad": int((both & (g["material"] != o["material"])).sum()),
    }
