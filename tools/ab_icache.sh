#!/bin/bash
# instruction-cache sensitivity A/B through the CLI host (C4 at 1000 spp): the product build, the opt-in chain instance, and
# build flavors preloaded over the product library (LD_PRELOAD=... as an environment argument)
B=./mu-lambda-raytracer_b200/rt_main
C4="--world=final_scene --seed=42 --aspect_ratio=1:1 --image_width=800 --samples_per_pixel=1000"
run() { for rep in 1 2; do env "$@" timeout 120 $B $C4 --stats 2>&1 >/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*', d['mpaths_per_s'], round(d['rays']/d['paths'],4))"; done; }
for e in "$@"; do run $e; done
