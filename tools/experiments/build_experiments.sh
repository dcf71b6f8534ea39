#!/bin/bash
# Rebuilds librt_b200.so WITH the rejected kernel placements of rt_persist.cu (RTB_PS_EXPERIMENTS: stack levels / scene
# copies in shared memory), for re-running the A/B of profiles/r2_ab_variants.txt:
#   bash tools/experiments/build_experiments.sh && gpurun -- 'bash tools/ab_r2.sh "RT_PS_VARIANT=1" ... "RT_PS_VARIANT=7"'
# `python -m mu_lambda_raytracer_b200.build --force` restores the product build.
set -e
cd "$(dirname "$0")/../.."
NVCC_APPEND_FLAGS="-DRTB_PS_EXPERIMENTS" python mu-lambda-raytracer_b200/build.py --force
