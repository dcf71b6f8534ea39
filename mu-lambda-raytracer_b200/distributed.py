"""Multi-GPU sharding of the render loop: one process per GPU (torch.distributed), sample slices, one reduce.

The reference parallelises over image rows on one host (rayon, src/raytrace.rs:176-185).  Here every (pixel, sample)
path is independent and keyed by its own Philox counter, so rank g of G renders ALL pixels for a disjoint run of
sample indices and the accumulation buffers are summed onto rank 0 with a single reduce (NCCL over NVLink on GPUs; the
same code runs on gloo/CPU tensors in the tests).  The buffers hold the library's fixed-point sums (int64 tensors,
2^-32 units): the reduce is an exact integer sum, so N ranks produce the very image one GPU produces.  The tonemap
(to_rgb, src/raytrace.rs:59-68) then runs on rank 0 only.
"""
import ctypes as C

from . import abi


def sample_slice(samples_per_pixel, world_size, rank, first_sample=0):
    """[begin, begin+count) of rank's samples: contiguous, disjoint, covering, sizes differ by at most one."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, extra = divmod(int(samples_per_pixel), world_size)
    count = base + (1 if rank < extra else 0)
    begin = first_sample + rank * base + min(rank, extra)
    return begin, count


def reduce_accumulation(accum, dst=0, group=None):
    """Sum the per-rank accumulation buffers onto rank `dst` (one collective; a no-op without a process group)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return accum


def render_sharded(scene, camera, width, height, samples_per_pixel, max_depth, seed, accum, rgb=None, first_sample=0,
                   pipeline=abi.RT_PIPELINE_AUTO, stream=None, group=None, stats=None):
    """Render this rank's sample slice into the CUDA tensor `accum` ([H, W, 3] int64 fixed-point sums, overwritten),
    reduce onto rank 0 and, on rank 0 with `rgb` ([H, W, 3] int32) given, tonemap.  Returns (begin, count) of the slice
    rendered here.  `stats`: an abi.RtStats to fill (synchronises the stream)."""
    import torch
    import torch.distributed as dist
    if accum.dtype != torch.int64:
        raise TypeError("accum must be an int64 tensor (fixed-point radiance sums)")
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    begin, count = sample_slice(samples_per_pixel, world, rank, first_sample)
    accum.zero_()
    lib = abi.load()
    s = stream if stream is not None else torch.cuda.current_stream(accum.device)
    if count > 0:
        p = abi.RtParams()
        p.width, p.height, p.samples_per_pixel, p.max_depth = width, height, samples_per_pixel, max_depth
        p.seed, p.sample_begin, p.sample_count, p.pipeline, p.device = seed, begin, count, pipeline, -1
        abi.check(lib.rt_render_accumulate_fixed_device(scene.handle, C.byref(camera.c), C.byref(p), accum.data_ptr(), s.cuda_stream,
                                                        C.byref(stats) if stats is not None else None))
    reduce_accumulation(accum, 0, group)
    if rank == 0 and rgb is not None:
        abi.check(lib.rt_tonemap_fixed_device(accum.data_ptr(), rgb.data_ptr(), width * height, samples_per_pixel, accum.device.index,
                                              s.cuda_stream))
    return begin, count
