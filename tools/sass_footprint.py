#!/usr/bin/env python
"""Static instruction footprint of one persist_kernel instance, by source function (nvdisasm line info of a -lineinfo build).
The kernel's speed hangs on the SM's 32 KB instruction cache, so this is the number to watch while editing.
usage: python tools/sass_footprint.py [kernel-name-substring]   (default: the product instance, bvh4 768 threads fast media)"""
import bisect
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "mu-lambda-raytracer_b200", "csrc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-ftz=true", "-prec-div=false", "-prec-sqrt=false"]


def main(want):
    tmp = tempfile.mkdtemp()
    cubin = os.path.join(tmp, "p.cubin")
    subprocess.check_call(["nvcc"] + FLAGS + sys.argv[2:] + ["-cubin", "-o", cubin, os.path.join(CSRC, "rt_persist.cu")])
    sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], stdout=subprocess.PIPE, text=True).stdout.split("\n")
    start = end = None
    for i, l in enumerate(sass):
        if start is None and l.startswith(".text.") and want in l:
            start = i
        elif start is not None and l.strip().startswith(".section"):
            end = i
            break
    sec = sass[start:end]
    dev = open(os.path.join(CSRC, "rt_device.cuh")).read().split("\n")
    funcs = [(i, m.group(1)) for i, l in enumerate(dev, 1) for m in [re.match(r"RTB_DEV(?:_NOINLINE)?\s+[\w:<>\*&\s]+?\s+\*?&?(\w+)\(", l)] if m]
    starts = [f[0] for f in funcs]
    cur, counts, body, total = None, Counter(), True, 0
    for l in sec:
        if ".type" in l and "@function" in l:
            body = False
        m = re.search(r'//## File "([^"]*)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
        if re.match(r"\s*/\*[0-9a-f]{4}\*/", l):
            total += 1
            if cur is None:
                key = "?"
            elif cur[0] == "rt_device.cuh":
                key = funcs[bisect.bisect_right(starts, cur[1]) - 1][1]
            elif cur[0] == "rt_persist.cu":
                key = "persist:%03d" % (cur[1] // 20 * 20)
            else:
                key = cur[0]
            counts[key] += 1
    print("%s: %d instructions = %.1f KB" % (want, total, total * 16 / 1024))
    for k, v in counts.most_common(60):
        print("%5d  %s" % (v, k))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "persist_kernelILb0ELb1ELi0ELi768ELi0ELi0ELi2ELb0")
