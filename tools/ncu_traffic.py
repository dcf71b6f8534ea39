#!/usr/bin/env python
"""DRAM and L2 bytes per camera path of one captured kernel launch -> profiles/<tag>_persist_traffic.json, the file
bench.py's roofline.traffic is read from (so that figure is regenerated with every capture instead of being a constant).
usage: python tools/ncu_traffic.py report.ncu-rep paths_of_the_captured_launch out.json"""
import csv
import io
import json
import subprocess
import sys


def main(rep, paths, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, vals = rows[0], rows[1], rows[2]
    col = {h: (u, v) for h, u, v in zip(head, units, vals)}

    def num(name):
        u, v = col[name]
        x = float(v.replace(",", ""))
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "sector": 1.0, "": 1.0}.get(u, 1.0)
        return x * scale

    dram = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    l2 = num("lts__t_sectors.sum") * 32.0
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], stdout=subprocess.PIPE, text=True).stdout.strip()
    d = {"report": rep, "commit": commit, "kernel": col.get("Kernel Name", ("", "?"))[1], "paths": int(paths),
         "duration_ms": num("gpu__time_duration.sum") / 1e6 if col["gpu__time_duration.sum"][0] in ("ns", "nsecond") else num("gpu__time_duration.sum"),
         "dram_bytes": dram, "l2_bytes": l2, "dram_bytes_per_path": dram / int(paths), "l2_bytes_per_path": l2 / int(paths),
         "l1_hit_pct": num("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": num("lts__t_sector_hit_rate.pct")}
    with open(out, "w") as f:
        json.dump(d, f, indent=1)
    print(json.dumps(d))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
