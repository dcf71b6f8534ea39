#!/bin/bash
# Mpaths/s of every pipeline on the BASELINE.json configs C1-C4 (reduced spp for the big ones), through the CLI host
B=./mu-lambda-raytracer_b200/rt_main
for pipe in persistent wavefront megakernel; do
  echo "C1 $pipe $($B --world=random --seed=42 --aspect_ratio=3:2 --image_width=400 --samples_per_pixel=50 --pipeline $pipe --stats 2>&1 >/dev/null | tail -1)"
  echo "C2 $pipe $($B --world=random --seed=42 --aspect_ratio=3:2 --image_width=1200 --samples_per_pixel=500 --aperture=0.1 --focus_dist=10.0 --pipeline $pipe --stats 2>&1 >/dev/null | tail -1)"
  echo "C3 $pipe $($B --world=cornell_smoke --aspect_ratio=1:1 --image_width=600 --samples_per_pixel=1000 --pipeline $pipe --stats 2>&1 >/dev/null | tail -1)"
  echo "C4/10 $pipe $($B --world=final_scene --seed=42 --aspect_ratio=1:1 --image_width=800 --samples_per_pixel=1000 --pipeline $pipe --stats 2>&1 >/dev/null | tail -1)"
done
