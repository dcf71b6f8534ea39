"""The callers either side of the path: the rt_main command line (src/main.rs re-hosted over the C ABI) and
rt_render_multi (sample slices over the GPUs of one process + one ncclReduce)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import mu_lambda_raytracer_b200 as rt
from mu_lambda_raytracer_b200 import abi
from mu_lambda_raytracer_b200.distributed import sample_slice
import support as S

CLI = os.path.join(S.ROOT, "mu-lambda-raytracer_b200", "rt_main")


def run_cli(*args):
    return subprocess.run([CLI, *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, cwd=S.ROOT)


def test_cli_rejects_like_clap_and_unwrap():
    assert os.path.exists(CLI), "rt_main is built by __graft_entry__.build()"
    r = run_cli("--world=nope")  # possible_values (main.rs:79-84) -> clap error, exit 1
    assert r.returncode == 1 and "isn't a valid value for '--world" in r.stderr and "final_scene" in r.stderr and r.stdout == ""
    assert run_cli("--frobnicate=1").returncode == 1
    assert run_cli("--image_width").returncode == 1           # value missing
    assert run_cli("--aspect_ratio=16x9").returncode == 101   # parse_aspect_ratio indexes v[1] -> panic (main.rs:49-52)
    assert run_cli("--samples_per_pixel=many").returncode == 101
    assert run_cli("--lookfrom=1,2").returncode == 101        # parse_vector (main.rs:54-62)
    assert run_cli("--seed=-4").returncode == 101             # u64
    h = run_cli("--help")
    assert h.returncode == 0
    for flag in ("aspect_ratio", "image_width", "samples_per_pixel", "max_depth", "lookfrom", "lookat", "up", "field_of_view", "aperture",
                 "focus_dist", "world", "seed", "randomized_rendering"):
        assert "--" + flag in h.stdout, flag
    assert "[default: 16:9]" in h.stdout and "[default: 400]" in h.stdout and "[default: 200]" in h.stdout and "[default: simple]" in h.stdout


def test_cli_without_a_device_fails_loudly():
    if abi.load().rt_device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = run_cli("--world=cornell_smoke", "--aspect_ratio=1:1", "--image_width=16", "--samples_per_pixel=1")
    assert r.returncode == 1 and "no CPU path" in r.stderr and not r.stdout.startswith("P3")


def test_sample_slice_abi_matches_python():
    lib = abi.load()
    for spp in (1, 7, 8, 50, 1000, 10000):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                b, c = C.c_int32(), C.c_int32()
                lib.rt_sample_slice(5, spp, world, r, C.byref(b), C.byref(c))
                assert (b.value, c.value) == sample_slice(spp, world, r, first_sample=5)
                cover.extend(range(b.value, b.value + c.value))
            assert cover == list(range(5, 5 + spp))


def test_render_multi_rejects_bad_arguments():
    lib = abi.load()
    assert lib.rt_render_multi(None, 0, None, None, None, None, abi.RtProgressFn(), None, None) == abi.RT_ERR_INVALID


def _parse_ppm(text):
    lines = text.split("\n")
    assert lines[0] == "P3" and lines[2] == "255" and lines[-1] == ""
    w, h = map(int, lines[1].split())
    body = np.array([list(map(int, ln.split())) for ln in lines[3:-1]], dtype=np.int32)
    assert body.shape == (w * h, 3)
    return body.reshape(h, w, 3)[::-1]  # back to row j = 0 at the bottom


@pytest.mark.gpu
def test_cli_ppm_matches_the_library_render():
    args = ["--world=cornell_smoke", "--seed=42", "--aspect_ratio=1:1", "--image_width=48", "--samples_per_pixel=16"]
    r = run_cli(*args, "--stats")
    assert r.returncode == 0, r.stderr
    assert "Rendered in " in r.stderr and "Done!" in r.stderr and '"paths": 36864' in r.stderr
    img = _parse_ppm(r.stdout)
    assert img.shape == (48, 48, 3) and img.min() >= 0 and img.max() <= 255
    world = rt.World("cornell_smoke")
    scene = rt.Scene(world.build(42))
    info = world.camera()
    cam = S.make_camera(info["lookfrom"], info["lookat"], info["field_of_view"], 1.0)
    ren = rt.Renderer.new_with_rng(cam, scene, world.background(), rt.RenderingParams(16, 48, 48), rt.RecursiveRayTracer(50), rt.SeedableRngator(42))
    rgb, _ = ren.render_arrays()
    # same seed -> same paths, and the sums are order-independent integers: the very same bytes (src/raytrace.rs:179,197)
    assert np.array_equal(rgb, img)
    r_again = run_cli(*args)
    assert r_again.returncode == 0 and r_again.stdout == r.stdout
    assert rt.to_ppm(img) == r.stdout
    # `--flag value` spelling, aperture/focus/lookfrom flags, 3:2 image height = (W / 1.5) as usize
    r2 = run_cli("--world", "random", "--seed", "42", "--aspect_ratio", "3:2", "--image_width", "50", "--samples_per_pixel", "2",
                 "--aperture", "0.1", "--focus_dist", "10.0", "--lookfrom", "13,2,3", "--max_depth", "8")
    assert r2.returncode == 0 and r2.stdout.startswith("P3\n50 33\n255\n")
    # without --seed two runs differ (thread_rng in the reference)
    a, b = run_cli(*args[:1], *args[2:]), run_cli(*args[:1], *args[2:])
    assert a.returncode == 0 and b.returncode == 0 and a.stdout != b.stdout


def _render_multi(scenes, cam, W, H, spp, seed):
    p = abi.RtParams()
    p.width, p.height, p.samples_per_pixel, p.max_depth, p.seed, p.device = W, H, spp, 50, seed, -1
    arr = (C.c_void_p * len(scenes))(*[s.handle for s in scenes])
    accum, rgb = np.empty((H, W, 3), np.float32), np.empty((H, W, 3), np.int32)
    st = abi.RtStats()
    abi.check(abi.load().rt_render_multi(arr, len(scenes), C.byref(cam.c), C.byref(p), accum.ctypes.data, rgb.ctypes.data, abi.RtProgressFn(), None,
                                         C.byref(st)))
    return accum, rgb, st


@pytest.mark.gpu
def test_render_multi_matches_single_device():
    world = rt.World("final_scene")
    desc = world.build(42)
    info = world.camera()
    cam = S.make_camera(info["lookfrom"], info["lookat"], info["field_of_view"], 1.0)
    W = H = 96
    spp = 20
    one = rt.Scene(desc, device=0)
    a1, rgb1, st1 = _render_multi([one], cam, W, H, spp, 5)
    assert st1.paths == W * H * spp
    n_dev = abi.load().rt_device_count()
    if n_dev < 2:
        one.close()
        pytest.skip("needs 2 GPUs for the sharded half of the test")
    scenes = [one] + [rt.Scene(desc, device=g) for g in range(1, min(n_dev, 4))]
    a2, rgb2, st2 = _render_multi(scenes, cam, W, H, spp, 5)
    # the same (pixel, sample) Philox streams, split by sample index across the devices: identical paths, and the
    # fixed-point sums are reduced with an exact integer sum -> identical images
    assert st2.paths == st1.paths and st2.rays == st1.rays
    assert np.array_equal(a1, a2) and np.array_equal(rgb1, rgb2)
    for s in scenes:
        s.close()
