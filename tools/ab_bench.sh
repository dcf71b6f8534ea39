#!/bin/bash
# A/B helper (run under gpurun): C4 throughput of the persistent pipeline, once with the defaults and once per
# argument, each argument being a space-separated list of VAR=value settings, e.g.
#   bash tools/ab_bench.sh "RT_PS_WORK=24" "RT_PS_STALL=10 RT_PS_LEAVE=4" "RT_BVH_MAX_LEAF=4" "RT_PS_CARVEOUT=50"
# Tunables: RT_PS_WORK / RT_PS_STALL / RT_PS_LEAVE / RT_PS_DESCEND (scheduling thresholds), RT_PS_BLOCKS_PER_SM,
# RT_PS_CARVEOUT (shared-memory carve-out in percent), RT_BVH_MAX_LEAF, RT_WF_SLOTS (global wavefront pool).
run() { env "$@" python bench.py --steps 3 --warmup 2 --spp 500 --pipeline ${PIPELINE:-persistent} --e2e-steps 0 --cpu-spp 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', round(d['value'],1))"; }
run DEFAULTS=1
for e in "$@"; do run $e; done
