#!/usr/bin/env python
"""Hottest source lines of a kernel from an .ncu-rep captured with --set full --import-source on (-lineinfo build).
usage: python tools/ncu_hot_lines.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys


def main(path, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file, head, lines = "?", None, []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif len(r) > 5 and r[0] == "Line No":
            head = r
        elif head and len(r) == len(head) and r[0] not in ("", "Line No"):
            d = dict(zip(head, r))
            num = lambda k: float(d.get(k, "0").replace(",", "") or 0) if d.get(k, "0") not in ("-", "") else 0.0
            lines.append((num("# Samples"), num("Instructions Executed"), num("Thread Instructions Executed"), cur_file, r[0], r[1].strip()[:110]))
    tot_s = sum(l[0] for l in lines) or 1
    tot_i = sum(l[1] for l in lines) or 1
    tot_t = sum(l[2] for l in lines) or 1
    print(f"{path}: {int(tot_s)} samples, {int(tot_i)} warp instructions, {tot_t / tot_i:.1f} threads/inst")
    print(f"{'samples%':>8} {'inst%':>6} {'thr/inst':>8}  file:line  source")
    for s, i, t, f, ln, src in sorted(lines, reverse=True)[:top]:
        print(f"{100 * s / tot_s:8.2f} {100 * i / tot_i:6.2f} {t / i if i else 0:8.1f}  {f}:{ln}  {src}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
