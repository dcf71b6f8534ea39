#!/usr/bin/env python
"""Instruction and stall-sample shares of a kernel by source region (groups of file:line ranges given on the command line
or, by default, the regions of rt_device.cuh as printed by `grep -n RTB_DEV`).  usage: ncu_groups.py report.ncu-rep"""
import csv, io, re, subprocess, sys, os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def device_regions():
    """function name -> (first line, last line) in rt_device.cuh, from the RTB_DEV definitions"""
    path = os.path.join(ROOT, "mu-lambda-raytracer_b200", "csrc", "rt_device.cuh")
    lines = open(path).read().split("\n")
    starts = []
    for i, l in enumerate(lines, 1):
        m = re.match(r"^(RTB_DEV_NOINLINE|RTB_DEV) .*?(\w+)\(", l)
        if m and i > 90:
            starts.append((i, m.group(2)))
    out = []
    for k, (i, name) in enumerate(starts):
        end = starts[k + 1][0] - 1 if k + 1 < len(starts) else len(lines)
        out.append((i, end, name))
    return out


GROUP_OF = {  # function -> coarse group
    "philox4x32_10v": "philox", "philox4x32_10": "philox", "rng_next4": "philox", "mulhi32": "philox", "u32_to_unit": "philox",
    "sample_unit_ball": "scatter", "scatter": "scatter", "material_color": "scatter", "scatter_event": "scatter", "background_color": "scatter",
    "surface_at": "surface_at", "generate_camera_ray": "camera", "wf_init_camera": "camera",
    "load_prim": "prim tests", "sphere_roots": "prim tests", "hit_sphere": "prim tests", "box_slabs": "prim tests", "hit_box": "prim tests", "hit_prim": "prim tests",
    "to_object_space": "prim tests", "sphere_needs_f64": "prim tests", "sphere_roots_f64": "f64 sphere", "prim_instance": "prim tests",
    "slab_node": "node slabs", "medium_interval": "media", "sample_media": "media", "sphere_interval": "media", "wf_presample_media": "media",
    "texture_leaf": "textures", "texture_value": "textures", "noise_value": "textures", "texture_needs_uv": "textures", "sphere_uv_v": "textures", "sphere_uv": "textures",
    "wf_shade_core": "shade glue", "f4": "shade glue",
}


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    regions = device_regions()
    cur_file, head = "?", None
    agg = {}
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif len(r) > 5 and r[0] == "Line No":
            head = r
        elif head and len(r) == len(head) and r[0] not in ("", "Line No"):
            d = dict(zip(head, r))
            num = lambda k: float(d.get(k, "0").replace(",", "") or 0) if d.get(k, "0") not in ("-", "") else 0.0
            ln = int(r[0])
            if cur_file == "rt_device.cuh":
                name = next((n for a, b, n in regions if a <= ln <= b), "vector/misc helpers" if ln < 120 else "?")
                g = GROUP_OF.get(name, name)
                if ln < 120:
                    g = "vector/misc helpers (shared)"
            else:
                g = cur_file + " (kernel body)"
            a = agg.setdefault(g, [0.0, 0.0, 0.0])
            a[0] += num("# Samples"); a[1] += num("Instructions Executed"); a[2] += num("Thread Instructions Executed")
    ts, ti = sum(a[0] for a in agg.values()) or 1, sum(a[1] for a in agg.values()) or 1
    print(f"{path}")
    print(f"{'samples%':>8} {'inst%':>6} {'thr/inst':>8}  region")
    for g, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{100 * a[0] / ts:8.2f} {100 * a[1] / ti:6.2f} {a[2] / a[1] if a[1] else 0:8.1f}  {g}")


if __name__ == "__main__":
    main(sys.argv[1])
