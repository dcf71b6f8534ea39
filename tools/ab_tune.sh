#!/bin/bash
# scheduling thresholds of the persistent kernel on C4 (1000 spp), one run each
B=./mu-lambda-raytracer_b200/rt_main
C4="--world=final_scene --seed=42 --aspect_ratio=1:1 --image_width=800 --samples_per_pixel=1000"
run() { env "$@" timeout 120 $B $C4 --stats 2>&1 >/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*', d['mpaths_per_s'])"; }
run X=0
for w in 22 24 26 30 32; do run RT_PS_WORK=$w; done
for s in 8 10 12 16 18 20; do run RT_PS_STALL=$s; done
for l in 4 6 10 12 16; do run RT_PS_LEAVE=$l; done
for d in 2 4 8 10 12; do run RT_PS_DESCEND=$d; done
run X=0
