#!/bin/bash
# throughput of the default build on C1-C4 through the CLI host (C4 at 1000 spp), with optional environment per run
B=./mu-lambda-raytracer_b200/rt_main
declare -A CFG
CFG[C1]="--world=random --seed=42 --aspect_ratio=3:2 --image_width=400 --samples_per_pixel=50"
CFG[C2]="--world=random --seed=42 --aspect_ratio=3:2 --image_width=1200 --samples_per_pixel=500 --aperture=0.1 --focus_dist=10.0"
CFG[C3]="--world=cornell_smoke --seed=42 --aspect_ratio=1:1 --image_width=600 --samples_per_pixel=1000"
CFG[C4]="--world=final_scene --seed=42 --aspect_ratio=1:1 --image_width=800 --samples_per_pixel=1000"
run() { cfg=$1; shift; for rep in 1 2; do env "$@" timeout 120 $B ${CFG[$cfg]} --stats 2>&1 >/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$cfg $*', d['mpaths_per_s'])"; done; }
for e in "${@:-X=0}"; do for c in C4 C3 C2 C1; do run $c $e; done; done
