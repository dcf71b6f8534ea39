#!/bin/bash
# feature-specialised kernel instances (RT_PS_FEAT): identical images against the generic instance, then throughput per config
B=./mu-lambda-raytracer_b200/rt_main
declare -A CFG
CFG[C1]="--world=random --seed=42 --aspect_ratio=3:2 --image_width=400 --samples_per_pixel=50"
CFG[C2]="--world=random --seed=42 --aspect_ratio=3:2 --image_width=1200 --samples_per_pixel=500 --aperture=0.1 --focus_dist=10.0"
CFG[C3]="--world=cornell_smoke --seed=42 --aspect_ratio=1:1 --image_width=600 --samples_per_pixel=1000"
CFG[C4]="--world=final_scene --seed=42 --aspect_ratio=1:1 --image_width=800 --samples_per_pixel=1000"
for w in "random 1.5" "cornell_smoke 1" "final_scene 1" "cornell_box 1" "earth 1.5" "simple_light 1.5" "random_chk 1.5"; do
  set -- $w
  A="--world=$1 --seed=42 --image_width=160 --samples_per_pixel=64"
  RT_PS_FEAT=0 timeout 120 $B $A > /tmp/gen.ppm 2>/dev/null; r1=$?
  timeout 120 $B $A > /tmp/feat.ppm 2>/dev/null; r2=$?
  echo "$1 rc=$r1/$r2 identical=$(cmp -s /tmp/gen.ppm /tmp/feat.ppm && echo yes || echo NO)"
done
run() { cfg=$1; shift; for rep in 1 2; do env "$@" timeout 120 $B ${CFG[$cfg]} --stats 2>&1 >/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$cfg $*', d['mpaths_per_s'])"; done; }
for c in C4 C3 C2 C1; do run $c RT_PS_FEAT=0; run $c RT_PS_FEAT=1; done
