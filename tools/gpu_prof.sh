#!/bin/bash
# one full ncu capture of one kernel of one pipeline: tools/gpu_prof.sh <pipeline> <kernel regex> <tag> [skip]
set -u
PIPE=$1; KREG=$2; TAG=$3; SKIP=${4:-1}
CMD="python bench.py --steps 1 --warmup 1 --spp 64 --pipeline $PIPE --e2e-steps 0 --cpu-spp 0"  # job of 64 spp: 40.96 M paths per launch
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$KREG -s $SKIP -c 1 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo prof_rc=$?
tail -1 gpurun_out/plain_$TAG.log | cut -c1-200
