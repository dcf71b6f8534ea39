#!/bin/bash
# small jobs (C1 and smaller) through the CLI host: 5 repetitions each, best and median of device ms
B=./mu-lambda-raytracer_b200/rt_main
run() { cfg=$1; shift; env "$@" python - "$cfg" "$*" <<'PY'
import subprocess, sys, json, os
cfg, label = sys.argv[1], sys.argv[2]
ms = []
for _ in range(7):
    out = subprocess.run(["./mu-lambda-raytracer_b200/rt_main"] + cfg.split() + ["--stats"], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True).stderr.strip().splitlines()[-1]
    d = json.loads(out); ms.append(d["device_ms"])
ms.sort()
print(label, "|", cfg.split()[0], cfg.split()[3], cfg.split()[4], "| device ms best %.3f median %.3f | Mpaths/s best %.0f" % (ms[0], ms[len(ms)//2], d["paths"] / ms[0] / 1e3))
PY
}
C1="--world=random --seed=42 --aspect_ratio=3:2 --image_width=400 --samples_per_pixel=50"
C1b="--world=final_scene --seed=42 --aspect_ratio=1:1 --image_width=400 --samples_per_pixel=50"
C1c="--world=cornell_smoke --seed=42 --aspect_ratio=1:1 --image_width=300 --samples_per_pixel=100"
for e in "$@"; do run "$C1" $e; run "$C1b" $e; run "$C1c" $e; done
