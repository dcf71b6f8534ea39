"""Round-2 contract of the C ABI, checked on the GPU: the per-row logger of Renderer::render (raytrace.rs:174,182), bit
reproducibility per seed (raytrace.rs:179,197), exact additivity of sample ranges in the fixed-point sums (what the
multi-GPU slicing relies on), the layout selector, the retired pipeline, cache release and the peaks probe."""
import ctypes as C

import numpy as np
import pytest

import mu_lambda_raytracer_b200 as rt
from mu_lambda_raytracer_b200 import abi
import support as S

pytestmark = pytest.mark.gpu


def _setup(name="cornell_smoke", aspect=1.0):
    world = rt.World(name)
    scene = rt.Scene(world.build(42))
    info = world.camera()
    cam = S.make_camera(info["lookfrom"], info["lookat"], info["field_of_view"], aspect)
    return world, scene, cam


def _params(W, H, spp, seed=42, begin=0, count=0, pipeline=abi.RT_PIPELINE_AUTO, layout=0):
    p = abi.RtParams()
    p.width, p.height, p.samples_per_pixel, p.max_depth, p.seed = W, H, spp, 50, seed
    p.sample_begin, p.sample_count, p.pipeline, p.device, p.bvh_layout = begin, count, pipeline, -1, layout
    return p


def test_logger_is_called_once_per_row_in_order():
    world, scene, cam = _setup()
    W, H = 96, 72
    rows = []
    cb = abi.RtProgressFn(lambda j, total, user: rows.append((j, total)))
    rgb = np.empty((H, W, 3), np.int32)
    abi.check(abi.load().rt_render(scene.handle, C.byref(cam.c), C.byref(_params(W, H, 8)), None, rgb.ctypes.data, cb, None, None))
    assert rows == [(j, H) for j in range(H)]  # logger(j, H), H times (raytrace.rs:182)
    # the Python mirror of Renderer::render passes it through
    seen = []
    r = rt.Renderer.new_with_rng(cam, scene, world.background(), rt.RenderingParams(4, H, W), rt.RecursiveRayTracer(50), rt.SeedableRngator(1))
    image = r.render(lambda j, total: seen.append(j))
    assert seen == list(range(H)) and len(image) == H and len(image[0]) == W and len(image[0][0]) == 3
    scene.close()


@pytest.mark.parametrize("pipeline", [abi.RT_PIPELINE_PERSISTENT, abi.RT_PIPELINE_MEGAKERNEL, abi.RT_PIPELINE_WAVEFRONT])
def test_one_seed_one_image_bit_for_bit(pipeline):
    """the reference is deterministic per seed; so is every pipeline here (order-independent fixed-point sums), run after run
    and however the sample range is cut"""
    world, scene, cam = _setup("final_scene")
    W = H = 200
    lib = abi.load()

    def render(begin, count, spp=48):
        accum, rgb = np.empty((H, W, 3), np.float32), np.empty((H, W, 3), np.int32)
        st = abi.RtStats()
        abi.check(lib.rt_render(scene.handle, C.byref(cam.c), C.byref(_params(W, H, spp, 7, begin, count, pipeline)), accum.ctypes.data, rgb.ctypes.data,
                                abi.RtProgressFn(), None, C.byref(st)))
        return accum, rgb, int(st.rays)

    a1, rgb1, rays1 = render(0, 48)
    a2, rgb2, rays2 = render(0, 48)
    assert rays1 == rays2 and np.array_equal(a1, a2) and np.array_equal(rgb1, rgb2)
    scene.close()


def test_sample_ranges_add_up_exactly_in_fixed_point():
    import torch
    world, scene, cam = _setup("final_scene")
    W = H = 160
    lib = abi.load()
    dev = torch.device("cuda", 0)

    def sums(begin, count):
        acc = torch.zeros(H, W, 3, dtype=torch.int64, device=dev)
        abi.check(lib.rt_render_accumulate_fixed_device(scene.handle, C.byref(cam.c), C.byref(_params(W, H, 64, 3, begin, count)), acc.data_ptr(), None, None))
        torch.cuda.synchronize()
        return acc

    whole = sums(0, 64)
    parts = sums(0, 10) + sums(10, 31) + sums(41, 23)
    assert torch.equal(whole, parts)  # integer sums: no rounding, any split of the sample range gives the same buffer
    # tonemap of the fixed-point sums == to_rgb of the float sums rt_render returns
    rgb = torch.zeros(H, W, 3, dtype=torch.int32, device=dev)
    abi.check(lib.rt_tonemap_fixed_device(whole.data_ptr(), rgb.data_ptr(), W * H, 64, 0, None))
    f = torch.zeros(H, W, 3, dtype=torch.float32, device=dev)
    abi.check(lib.rt_accum_fixed_to_float_device(whole.data_ptr(), f.data_ptr(), 3 * W * H, 0, None))
    torch.cuda.synchronize()
    accum, rgb_host = np.empty((H, W, 3), np.float32), np.empty((H, W, 3), np.int32)
    abi.check(lib.rt_render(scene.handle, C.byref(cam.c), C.byref(_params(W, H, 64, 3)), accum.ctypes.data, rgb_host.ctypes.data, abi.RtProgressFn(), None, None))
    assert np.array_equal(rgb.cpu().numpy(), rgb_host) and np.array_equal(f.cpu().numpy(), accum)
    assert np.array_equal(accum, (whole.cpu().numpy().astype(np.float64) / abi.RT_ACCUM_FIXED_ONE).astype(np.float32))
    scene.close()


def test_layout_selector_and_retired_pipeline():
    world, scene, cam = _setup()
    lib = abi.load()
    W = H = 64
    rgb = np.empty((H, W, 3), np.int32)
    images = {}
    for layout in (0, 2, 4):
        st = abi.RtStats()
        abi.check(lib.rt_render(scene.handle, C.byref(cam.c), C.byref(_params(W, H, 16, layout=layout)), None, rgb.ctypes.data, abi.RtProgressFn(), None, C.byref(st)))
        assert st.pipeline_used == abi.RT_PIPELINE_PERSISTENT and st.bvh_layout_used == (layout or 4)
        images[layout] = rgb.copy()
    # cornell_smoke has no two primitives at the same distance along a ray: both tree layouts find the same hits
    assert np.array_equal(images[2], images[4]) and np.array_equal(images[0], images[4])
    assert lib.rt_render(scene.handle, C.byref(cam.c), C.byref(_params(W, H, 16, layout=3)), None, rgb.ctypes.data, abi.RtProgressFn(), None, None) == abi.RT_ERR_INVALID
    rc = lib.rt_render(scene.handle, C.byref(cam.c), C.byref(_params(W, H, 16, pipeline=abi.RT_PIPELINE_WAVEFRONT_SMEM)), None, rgb.ctypes.data, abi.RtProgressFn(), None, None)
    assert rc == abi.RT_ERR_UNSUPPORTED and b"tools/experiments" in lib.rt_last_error()
    scene.close()


def test_cached_memory_can_be_released():
    import torch
    lib = abi.load()
    world, scene, cam = _setup("final_scene")
    rgb = np.empty((128, 128, 3), np.int32)
    abi.check(lib.rt_render(scene.handle, C.byref(cam.c), C.byref(_params(128, 128, 4)), None, rgb.ctypes.data, abi.RtProgressFn(), None, None))
    scene.close()  # blocks go to the library's free list ...
    free_before, _ = torch.cuda.mem_get_info(0)
    lib.rt_release_cached_memory()  # ... and back to the driver here
    free_after, _ = torch.cuda.mem_get_info(0)
    assert free_after >= free_before + (2 << 20)  # at least the 2 MB earth texture and the scratch images
    world, scene, cam = _setup("final_scene")  # and the library keeps working
    rgb2 = np.empty((128, 128, 3), np.int32)
    abi.check(lib.rt_render(scene.handle, C.byref(cam.c), C.byref(_params(128, 128, 4)), None, rgb2.ctypes.data, abi.RtProgressFn(), None, None))
    assert np.array_equal(rgb, rgb2)
    scene.close()


def test_measured_peaks_are_sane():
    pk = abi.RtPeaks()
    abi.check(abi.load().rt_measure_peaks(0, C.byref(pk)))
    assert pk.sm_count >= 100 and 30.0 < pk.fp32_ffma_tflops <= 1.02 * pk.fp32_theoretical_tflops
    assert 30.0 < pk.fp32_ffma2_tflops <= 1.02 * pk.fp32_theoretical_tflops
    assert 3000.0 < pk.l2_read_gbs < 40000.0
