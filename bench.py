#!/usr/bin/env python
"""Headline benchmark: Mpaths/s (camera paths per second) on final_scene 800x800, seed 42 (BASELINE.json C4).

  python bench.py --gpus N --steps K --warmup W            this repo's CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W    the reference algorithm on the host CPU cores

A "step" is one pass of the hot path over one batch of synthetic work: the full 800x800 frame at `--spp` samples
per pixel on every GPU (weak scaling: per-GPU work is fixed), followed — inside the timed region — by the NCCL
reduce of the accumulation buffers onto rank 0 and the tonemap.  The C4 job (10 000 spp) is 10 such steps at the
default 1000 spp; cost is linear in spp because samples are independent (src/raytrace.rs:190-195).

The reference is a Rust program and there is no Rust toolchain in this image, so the reference arm times the
oracle — the C++ f64 restatement of the reference's algorithm (oracle/) — on all host cores.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

WORLD, SEED, WIDTH, HEIGHT, ASPECT, MAX_DEPTH, JOB_SPP = "final_scene", 42, 800, 800, 1.0, 50, 10000


def baseline_metric():
    """the headline metric exactly as BASELINE.json words it (unit: Mpaths/s = million camera paths per second)"""
    try:
        with open(os.path.join(ROOT, "BASELINE.json")) as f:
            return json.load(f)["metric"]
    except Exception:
        return "Mpaths/s on final_scene 800x800 (device-timed) vs host-CPU reference"

# Algorithmic work of the REFERENCE's traversal per camera path on C4, counted by the oracle's instrumentation
# (tests/golden/algo_work.py: 800x800, 16 spp, seed 42).  flops = 27*aabb + 45*sphere + 15*rect + 12*xform
# + 40*medium + shade terms; bytes = 32 B per node / primitive record touched.
# DRAM bytes per camera path of the dominant kernel, from the committed ncu --set full captures (profiles/): C4 at 64 spp,
# dram__bytes_read.sum + dram__bytes_write.sum of one launch divided by the paths of that launch
NCU_DRAM_BYTES_PER_PATH = {"persistent": 18.72e6 / 40.96e6, "megakernel": 10.56e6 / 40.96e6, "wavefront": 597e6 / 504e3}

ALGO = {"rays_per_path": 4.159, "aabb": 66.59, "sphere": 42.99, "rect": 60.77, "xform": 8.318, "medium": 8.318,
        "lambertian": 0.878, "metal": 0.0312, "dielectric": 0.315, "isotropic": 1.963, "perlin": 0.0682, "image": 0.0695,
        "background": 0.886}  # tests/golden/algo_work.py 16 (frozen in BASELINE.md section 4)


def algo_flops_per_path():
    a = ALGO
    geometry = 27 * a["aabb"] + 45 * a["sphere"] + 15 * a["rect"] + 12 * a["xform"] + 40 * a["medium"]
    shade = (40 * a["lambertian"] + 60 * a["metal"] + 80 * a["dielectric"] + 25 * a["isotropic"] + 1400 * a["perlin"]
             + 55 * a["image"] + 20 * a["background"])
    return geometry + shade


def algo_bytes_per_path():
    a = ALGO
    return 32 * (a["aabb"] + a["sphere"] + a["rect"]) + 48 * a["medium"]


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(s) > 3 + k and s[3 + k].lower().startswith("active") for s in self.samples)]
        power = [float(s[2]) for s in self.samples if s[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "power_w_max": max(power) if power else None, "samples": len(self.samples), "reasons": reasons}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def cpu_reference_run(spp, steps, warmup, threads=0):
    """Times the oracle (reference algorithm, f64, PCG64 row streams, row-parallel) on the full 800x800 frame at `spp`
    samples per step.  Returns (Mpaths/s, ms per step, cores, rays per path)."""
    import support as S
    ow = S.OracleWorld(WORLD, SEED)
    cam = S.make_camera(ow.lookfrom, ow.lookat, ow.vfov, ASPECT)
    cores = threads or S.oracle().orc_hardware_threads()
    for w in range(warmup):
        ow.render(cam.c, WIDTH, HEIGHT, 1, MAX_DEPTH, render_seed=SEED + 100000 * (w + 1), threads=cores, rows=(0, 80))
    secs, rays, paths = 0.0, 0, 0
    for k in range(steps):
        _, _, counters, s = ow.render(cam.c, WIDTH, HEIGHT, spp, MAX_DEPTH, render_seed=SEED + 1000 * k, threads=cores)
        secs += s
        paths += int(counters[0])
        rays += int(counters[1])
    return paths / secs / 1e6, 1e3 * secs / steps, cores, rays / max(paths, 1)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    spp = args.ref_spp
    value, ms, cores, rpp = cpu_reference_run(spp, args.steps, args.warmup)
    sample = f"full 800x800 frame at {spp} spp per step ({WIDTH * HEIGHT * spp} camera paths; the C4 job is {JOB_SPP} spp), max_depth {MAX_DEPTH}"
    line = {
        "impl": "reference", "metric": baseline_metric(), "value": value, "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"C4 {WORLD} {WIDTH}x{HEIGHT} seed {SEED} max_depth {MAX_DEPTH}", "spp_per_step": spp,
                   "note": "reference = C++ f64 restatement of the Rust renderer (oracle/); no Rust toolchain in this image"},
        "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "rays_per_path": rpp, "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def run_b200(args):
    import torch
    import mu_lambda_raytracer_b200 as rt
    from mu_lambda_raytracer_b200 import abi

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the render path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world_size > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = abi.load()
    dev = torch.device("cuda", local_rank)

    world = rt.World(WORLD)
    desc = world.build(SEED)
    scene = rt.Scene(desc, device=local_rank)
    info = world.camera()
    focus = float(np.linalg.norm(np.asarray(info["lookat"]) - np.asarray(info["lookfrom"])))
    cam = rt.Camera(info["lookfrom"], info["lookat"], (0, 1, 0), info["field_of_view"], ASPECT, 0.0, focus)

    spp = args.spp
    total_spp = spp * world_size  # the image every step produces has spp samples from each GPU
    pipeline = {"auto": abi.RT_PIPELINE_AUTO, "megakernel": abi.RT_PIPELINE_MEGAKERNEL, "wavefront": abi.RT_PIPELINE_WAVEFRONT,
                "wavefront_smem": abi.RT_PIPELINE_WAVEFRONT_SMEM, "persistent": abi.RT_PIPELINE_PERSISTENT}[args.pipeline]

    def params(step):
        p = abi.RtParams()
        p.width, p.height, p.samples_per_pixel, p.max_depth = WIDTH, HEIGHT, total_spp, MAX_DEPTH
        p.seed = SEED
        # disjoint Philox sample indices per rank and per step
        p.sample_begin, p.sample_count = (step * world_size + rank) * spp, spp
        assert p.sample_begin + spp < 2 ** 31
        p.pipeline, p.device, p.samples_per_item = pipeline, -1, args.samples_per_item
        return p

    accum = torch.zeros(HEIGHT, WIDTH, 3, dtype=torch.float32, device=dev)
    rgb = torch.zeros(HEIGHT, WIDTH, 3, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream()
    launches = [0]
    rays = [0]

    def step(i, count_stats=False):
        flush.fill_(i & 0xFF)  # L2 flush between steps
        accum.zero_()
        p = params(i)
        st = abi.RtStats() if count_stats else None
        abi.check(lib.rt_render_accumulate_device(scene.handle, C.byref(cam.c), C.byref(p), accum.data_ptr(), stream.cuda_stream,
                                                  C.byref(st) if st else None))
        if dist is not None:
            dist.reduce(accum, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            abi.check(lib.rt_tonemap_device(accum.data_ptr(), rgb.data_ptr(), WIDTH * HEIGHT, total_spp, local_rank, stream.cuda_stream))
        if st:
            launches[0] = st.kernel_launches + (1 if rank == 0 else 0)
            rays[0] = st.rays

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(args.warmup):
        step(1000 + w, count_stats=(w == 0))
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        step(k)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.summary() if sampler else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    paths_per_step = WIDTH * HEIGHT * spp * world_size
    value = paths_per_step * args.steps / (ms_total / 1e3) / 1e6

    # ---- the render region alone (CUDA events recorded by the library on the launching stream around ALL kernels of the
    # pipeline; no flush / reduce / tonemap): the numerator of the roofline figure
    p = params(2000)
    st = abi.RtStats()
    accum.zero_()
    torch.cuda.synchronize()
    abi.check(lib.rt_render_accumulate_device(scene.handle, C.byref(cam.c), C.byref(p), accum.data_ptr(), stream.cuda_stream, C.byref(st)))
    render_ms = st.device_ms
    rays_per_path = st.rays / max(st.paths, 1)
    pipeline_used = {abi.RT_PIPELINE_MEGAKERNEL: "megakernel", abi.RT_PIPELINE_WAVEFRONT: "wavefront",
                     abi.RT_PIPELINE_WAVEFRONT_SMEM: "wavefront_smem", abi.RT_PIPELINE_PERSISTENT: "persistent"}.get(st.pipeline_used, str(st.pipeline_used))
    kernels = {"megakernel": "render_items_kernel", "wavefront": "wf_extend_kernel + wf_shade_kernel (one pair per round)",
               "wavefront_smem": "warpfront_kernel", "persistent": "persist_kernel"}.get(pipeline_used, "?")

    # ---- end to end through the public host API with HOST buffers: scene upload + render + readback, every step
    e2e = None
    if args.e2e_steps > 0:
        scene_bytes = scene.info()["device_bytes"]
        host_rgb = torch.empty(HEIGHT, WIDTH, 3, dtype=torch.int32).pin_memory()
        t0 = 0.0
        for k in range(-1, args.e2e_steps):  # k = -1: one untimed warm-up step (first-use allocations of the library)
            if k == 0:
                barrier()
                t0 = time.perf_counter()
            sc = rt.Scene(desc, device=local_rank)  # rt_scene_create: flatten + BVH build + H2D upload of the scene
            pk = params(3001 + k)
            if dist is None:
                # one GPU: the reference-facing call itself, Renderer::render with HOST buffers (rt_render)
                abi.check(lib.rt_render(sc.handle, C.byref(cam.c), C.byref(pk), None, host_rgb.data_ptr(), abi.RtProgressFn(), None, None))
            else:
                abi.check(lib.rt_render_accumulate_device(sc.handle, C.byref(cam.c), C.byref(pk), accum.zero_().data_ptr(), stream.cuda_stream, None))
                dist.reduce(accum, dst=0, op=dist.ReduceOp.SUM)
                if rank == 0:
                    abi.check(lib.rt_tonemap_device(accum.data_ptr(), rgb.data_ptr(), WIDTH * HEIGHT, total_spp, local_rank, stream.cuda_stream))
                    host_rgb.copy_(rgb, non_blocking=True)
                torch.cuda.synchronize()
            sc.close()
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": paths_per_step * args.e2e_steps / float(tt.item()) / 1e6, "unit": "Mpaths/s",
               "h2d_bytes_per_step": int(scene_bytes) * world_size, "d2h_bytes_per_step": int(host_rgb.numel() * 4),
               "api": ("rt_scene_create [flatten+BVH+upload] -> rt_render(host rgb buffer) -> rt_scene_destroy" if dist is None else
                       "rt_scene_create -> rt_render_accumulate_device -> ncclReduce -> rt_tonemap_device -> pinned host rgb"),
               "steps": args.e2e_steps}

    cpu = None
    if rank == 0 and world_size == 1 and args.cpu_spp > 0:
        v, ms, cores, rpp = cpu_reference_run(args.cpu_spp, 1, 1)
        cpu = {"value": v, "unit": "Mpaths/s", "cores": cores, "kind": "port",
               "sample": f"full 800x800 frame at {args.cpu_spp} spp ({WIDTH * HEIGHT * args.cpu_spp} paths, {ms / 1e3:.1f} s), C++ f64 restatement of the reference (no Rust toolchain)",
               "rays_per_path": rpp}

    if rank == 0:
        peaks, how = measured_peaks()
        n_sm = torch.cuda.get_device_properties(local_rank).multi_processor_count
        fp32_peak = n_sm * 128 * 2 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12
        flops = algo_flops_per_path() * WIDTH * HEIGHT * spp
        achieved = flops / (render_ms / 1e3) / 1e12
        accum_bytes = 3 * 4 * WIDTH * HEIGHT * spp  # one float RED per channel per terminated path (upper bound)
        line = {
            "metric": baseline_metric(), "value": value, "unit": "Mpaths/s", "n_gpus": world_size, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"C4 {WORLD} {WIDTH}x{HEIGHT} seed {SEED} max_depth {MAX_DEPTH}", "spp_per_step_per_gpu": spp,
                       "paths_per_step": paths_per_step, "job_spp": JOB_SPP, "pipeline": pipeline_used,
                       "parallelism": f"sample-slices x{world_size} + ncclReduce(sum) of the fp32 accumulation buffer",
                       "l2": "256 MB flush write between steps (inside the timed region)"},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": int(launches[0]) * args.steps,
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                         "traffic": (NCU_DRAM_BYTES_PER_PATH[pipeline_used] * WIDTH * HEIGHT * spp if pipeline_used in NCU_DRAM_BYTES_PER_PATH else None),
                         "traffic_note": "bytes per step = ncu dram bytes per path (profiles/, 64-spp capture) x paths per step",
                         "kernel": kernels, "kernel_ms": render_ms, "kernel_launches": st.kernel_launches,
                         "paths_per_kernel_ms": WIDTH * HEIGHT * spp,
                         "algorithmic_flops_per_path": algo_flops_per_path(), "algorithmic_bytes_per_path": algo_bytes_per_path(),
                         "peak_source": f"{n_sm} SMs x 128 lanes x 2 x sm_max_mhz from MEASURED_PEAKS.json ({how}); no measured FP32 peak exists",
                         "hbm": {"achieved_gbs": accum_bytes / (render_ms / 1e3) / 1e9, "peak_gbs": peaks.get("hbm_gbs"),
                                 "note": "algorithmic HBM bytes = accumulation-buffer REDs only; the scene (~100 KB + 2 MB texture) is L1/L2 resident"}},
            "cpu_baseline": cpu,
            "rays_per_path": rays_per_path, "mrays_per_s": value * rays_per_path,
        }
        print(json.dumps(line), flush=True)
    scene.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--spp", type=int, default=1000, help="samples per pixel per step per GPU")
    ap.add_argument("--pipeline", default="auto", choices=["auto", "megakernel", "wavefront", "wavefront_smem", "persistent"])
    ap.add_argument("--samples-per-item", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-spp", type=int, default=128, help="spp of the bounded CPU-baseline sample (0 = skip)")
    ap.add_argument("--ref-spp", type=int, default=32, help="spp per step of --impl reference")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
