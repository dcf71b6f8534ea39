#!/bin/bash
run() { env "$@" python bench.py --steps 3 --warmup 2 --spp 500 --pipeline persistent --e2e-steps 0 --cpu-spp 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', round(d['value'],1))"; }
run RT_BVH_MAX_LEAF=1
run RT_BVH_MAX_LEAF=1 RT_PS_WORK=28
run RT_BVH_MAX_LEAF=1 RT_PS_WORK=32
run RT_BVH_MAX_LEAF=1 RT_PS_WORK=32 RT_PS_STALL=8
run RT_BVH_MAX_LEAF=1 RT_PS_WORK=28 RT_PS_STALL=20
run RT_BVH_MAX_LEAF=1 RT_PS_WORK=30 RT_PS_STALL=4
B=./mu-lambda-raytracer_b200/rt_main
for l in 1 4; do
  echo "leaf=$l C2 $(RT_BVH_MAX_LEAF=$l $B --world=random --seed=42 --aspect_ratio=3:2 --image_width=1200 --samples_per_pixel=500 --aperture=0.1 --focus_dist=10.0 --stats 2>&1 >/dev/null | tail -1 | grep -o 'mpaths.*')"
  echo "leaf=$l C3 $(RT_BVH_MAX_LEAF=$l $B --world=cornell_smoke --aspect_ratio=1:1 --image_width=600 --samples_per_pixel=1000 --stats 2>&1 >/dev/null | tail -1 | grep -o 'mpaths.*')"
done
