#!/bin/bash
# A/B through the CLI host on one BASELINE config: tools/ab_cli.sh C2 "RT_BVH_LAYOUT=2" "RT_PS_VARIANT=1" ...
B=./mu-lambda-raytracer_b200/rt_main
case $1 in
  C1) ARGS="--world=random --seed=42 --aspect_ratio=3:2 --image_width=400 --samples_per_pixel=50";;
  C2) ARGS="--world=random --seed=42 --aspect_ratio=3:2 --image_width=1200 --samples_per_pixel=500 --aperture=0.1 --focus_dist=10.0";;
  C3) ARGS="--world=cornell_smoke --aspect_ratio=1:1 --image_width=600 --samples_per_pixel=1000";;
  C4) ARGS="--world=final_scene --seed=42 --aspect_ratio=1:1 --image_width=800 --samples_per_pixel=1000";;
esac
shift
for e in "$@"; do
  for rep in 1 2; do env $e $B $ARGS --stats 2>&1 >/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$e', d['mpaths_per_s'])"; done
done
