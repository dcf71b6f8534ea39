"""Image parity AT the BASELINE.json configurations: C1 exactly as quoted (400x266, 50 spp), C2, C3 and
C4 at their full resolutions with the sample count reduced to what the oracle — the CPU restatement of the reference —
finishes in seconds (cost is linear in spp, raytrace.rs:190-195; the full-spp C2/C4 images are checked against the
reference's published renders in test_published_renders.py).  Criterion of BASELINE.json north_star: RMSE(GPU, oracle)
<= 1.1 x RMSE(oracle seed A, oracle seed B) on display values, plus per-channel mean radiance within 1 %."""
import numpy as np
import pytest

import mu_lambda_raytracer_b200 as rt
from mu_lambda_raytracer_b200 import abi
import support as S

pytestmark = pytest.mark.gpu

CONFIGS = {
    # name: world, W, H, aspect flag, spp, aperture, focus_dist
    "C1": ("random", 400, 266, 1.5, 50, 0.0, None),           # --image_width=400 --samples_per_pixel=50: in full
    "C2": ("random", 1200, 800, 1.5, 24, 0.1, 10.0),          # 500 spp in the config
    "C3": ("cornell_smoke", 600, 600, 1.0, 48, 0.0, None),    # 1000 spp in the config
    "C4": ("final_scene", 800, 800, 1.0, 24, 0.0, None),      # 10 000 spp in the config
}


@pytest.mark.parametrize("cfg", list(CONFIGS))
@pytest.mark.parametrize("layout", [4, 2])
def test_config_image_within_noise_floor(cfg, layout):
    name, W, H, aspect, spp, aperture, focus = CONFIGS[cfg]
    assert H == int(W / aspect)  # main.rs:121
    world = rt.World(name)
    scene = rt.Scene(world.build(42))
    ow = S.OracleWorld(name, 42)
    if focus is None:
        focus = float(np.linalg.norm(np.asarray(ow.lookat) - np.asarray(ow.lookfrom)))
    cam = S.make_camera(ow.lookfrom, ow.lookat, ow.vfov, aspect, aperture=aperture, focus_dist=focus)
    r = rt.Renderer.new_with_rng(cam, scene, world.background(), rt.RenderingParams(spp, H, W), rt.RecursiveRayTracer(50), rt.SeedableRngator(42))
    r.bvh_layout = layout
    rgb, accum = r.render_arrays()
    assert r.stats["paths"] == W * H * spp and r.stats["bvh_layout"] == layout
    a1, _, c1, _ = ow.render(cam.c, W, H, spp, render_seed=42)
    a2, _, _, _ = ow.render(cam.c, W, H, spp, render_seed=977)
    rm = lambda x, y: float(np.sqrt(np.mean((x - y) ** 2)))
    disp = lambda a: np.sqrt(np.clip(a / spp, 0.0, 1.0))
    g = accum.astype(np.float64)
    floor = rm(disp(a1), disp(a2))
    got = 0.5 * (rm(disp(g), disp(a1)) + rm(disp(g), disp(a2)))
    assert got <= 1.1 * floor, f"{cfg}: RMSE {got:.5f} vs noise floor {floor:.5f}"
    for c in range(3):  # no colour-dependent bias: per-channel mean radiance within 1 % (+ the oracle's own seed-to-seed spread)
        m1, m2, mg = a1[..., c].mean(), a2[..., c].mean(), g[..., c].mean()
        assert abs(mg - 0.5 * (m1 + m2)) <= 0.01 * m1 + 3 * abs(m1 - m2), (cfg, c, mg, m1, m2)
    assert abs(r.stats["rays"] / r.stats["paths"] - c1[1] / c1[0]) < 0.02 * c1[1] / c1[0]
    scene.close()
