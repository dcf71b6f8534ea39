#!/bin/bash
# ncu captures of persist_kernel through the CLI host (C4 at 64 spp), one per environment given as arguments: tools/prof_chain.sh tag "ENV=.." ...
B=./mu-lambda-raytracer_b200/rt_main
ARGS="--world=final_scene --seed=42 --aspect_ratio=1:1 --image_width=800 --samples_per_pixel=64"
i=0
for e in "$@"; do
  i=$((i+1))
  env $e timeout 300 ncu --set full --clock-control none --import-source on -k regex:persist_kernel -c 1 -f -o gpurun_out/prof_chain_$i $B $ARGS > /dev/null 2> gpurun_out/prof_chain_$i.log; echo "$e rc=$?"
done
