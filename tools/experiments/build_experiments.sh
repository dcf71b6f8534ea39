#!/bin/bash
# Rebuilds librt_b200.so WITH the rejected kernel placements of rt_persist.cu (RTB_PS_EXPERIMENTS: stack levels / scene
# copies in shared memory), for re-running the A/B of profiles/r2_ab_variants.txt:
#   bash tools/experiments/build_experiments.sh && gpurun -- 'L=$PWD/mu-lambda-raytracer_b200/librt_b200_exp.so; bash tools/ab_r2.sh "RT_B200_LIB=$L RT_PS_VARIANT=2" ... "RT_B200_LIB=$L RT_PS_VARIANT=10"'
# The product library is not touched.
set -e
cd "$(dirname "$0")/../.."
RT_BUILD_FLAVOR=exp python mu-lambda-raytracer_b200/build.py
echo "built mu-lambda-raytracer_b200/librt_b200_exp.so: select it with RT_B200_LIB=... and a kernel instance with RT_PS_VARIANT=k"
