#!/bin/bash
# how far the feature-specialised instances are from the generic one (same paths up to FMA contraction: a few pixels by an LSB)
B=./mu-lambda-raytracer_b200/rt_main
for w in random final_scene earth simple_light two_spheres; do
  A="--world=$w --seed=42 --image_width=200 --samples_per_pixel=100"
  RT_PS_FEAT=0 timeout 120 $B $A --stats > /tmp/gen.ppm 2>/tmp/gen.err; timeout 120 $B $A --stats > /tmp/feat.ppm 2>/tmp/feat.err
  python - "$w" <<'PY'
import sys, json, numpy as np
def load(p):
    t = open(p).read().split()
    return np.array(t[4:], dtype=np.int64)
a, b = load("/tmp/gen.ppm"), load("/tmp/feat.ppm")
ra = json.loads(open("/tmp/gen.err").read().strip().splitlines()[-1]); rb = json.loads(open("/tmp/feat.err").read().strip().splitlines()[-1])
print(sys.argv[1], "values", a.size, "differ", int((a != b).sum()), "max", int(np.abs(a - b).max()), "rays", ra["rays"], rb["rays"])
PY
done
