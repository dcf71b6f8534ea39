"""The N > 1 host logic on CPU: world_size-2 gloo process group, sample-slice sharding and the single reduce.
Each rank integrates its slice with the host build of the device code (tests/emul) into the library's fixed-point sums;
the reduced buffer must equal — exactly — what one process renders for the whole sample range, and the slices must
tile the range."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sample_slices_tile_the_range():
    from mu_lambda_raytracer_b200.distributed import sample_slice
    for spp in (1, 7, 8, 50, 1000, 10000):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                b, c = sample_slice(spp, world, r, first_sample=5)
                cover.extend(range(b, b + c))
                assert c in (spp // world, spp // world + 1)
            assert cover == list(range(5, 5 + spp))
    with pytest.raises(ValueError):
        sample_slice(10, 2, 2)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import mu_lambda_raytracer_b200 as rt
    from mu_lambda_raytracer_b200.distributed import reduce_accumulation, sample_slice
    import support as S
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        world_obj = rt.World("cornell_smoke")
        desc = world_obj.build(42)
        es = S.EmulScene(desc.ptr)
        info = world_obj.camera()
        cam = S.make_camera(info["lookfrom"], info["lookat"], info["field_of_view"], 1.0)
        W = H = 24
        spp = 9
        begin, count = sample_slice(spp, world, rank)
        part, _ = es.render(cam, W, H, count, seed=11, threads=1, sample_begin=begin, fixed=True)
        np.save(os.path.join(out_dir, f"part{rank}.npy"), part)
        t = torch.from_numpy(part.view(np.int64).copy())  # what distributed.render_sharded reduces: int64 sums
        reduce_accumulation(t, dst=0)
        if rank == 0:
            np.save(os.path.join(out_dir, "reduced.npy"), t.numpy().view(np.uint64))
    finally:
        dist.destroy_process_group()


def test_two_rank_reduce_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    p0, p1 = np.load(tmp_path / "part0.npy"), np.load(tmp_path / "part1.npy")
    red = np.load(tmp_path / "reduced.npy")
    assert np.array_equal(red, p0 + p1)  # integer sums: exact
    # and the two slices together are the same paths a single process would have traced
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import mu_lambda_raytracer_b200 as rt
    import support as S
    world_obj = rt.World("cornell_smoke")
    desc = world_obj.build(42)
    es = S.EmulScene(desc.ptr)
    info = world_obj.camera()
    cam = S.make_camera(info["lookfrom"], info["lookat"], info["field_of_view"], 1.0)
    full, _ = es.render(cam, 24, 24, 9, seed=11, threads=1, fixed=True)
    assert np.array_equal(full, red)  # fixed-point accumulation: the split over ranks does not change a single bit
