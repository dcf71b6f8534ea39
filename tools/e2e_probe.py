import sys, time, ctypes as C
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import mu_lambda_raytracer_b200 as rt
from mu_lambda_raytracer_b200 import abi
lib = abi.load()
torch.cuda.set_device(0)
world = rt.World("final_scene"); desc = world.build(42)
info = world.camera()
focus = float(np.linalg.norm(np.asarray(info["lookat"]) - np.asarray(info["lookfrom"])))
cam = rt.Camera(info["lookfrom"], info["lookat"], (0, 1, 0), info["field_of_view"], 1.0, 0.0, focus)
host_rgb = torch.empty(800, 800, 3, dtype=torch.int32).pin_memory()
for k in range(8):
    p = abi.RtParams(); p.width = p.height = 800; p.samples_per_pixel = 1000; p.max_depth = 50; p.seed = 42
    p.sample_begin, p.sample_count, p.pipeline, p.device = 0, 1000, 0, -1
    t0 = time.perf_counter(); sc = rt.Scene(desc, device=0); t1 = time.perf_counter()
    st = abi.RtStats()
    abi.check(lib.rt_render(sc.handle, C.byref(cam.c), C.byref(p), None, host_rgb.data_ptr(), abi.RtProgressFn(), None, C.byref(st)))
    t2 = time.perf_counter(); sc.close(); t3 = time.perf_counter()
    print(f"create {1e3*(t1-t0):7.1f} ms  render {1e3*(t2-t1):7.1f} ms (device {st.device_ms:7.1f})  destroy {1e3*(t3-t2):7.1f} ms", flush=True)
