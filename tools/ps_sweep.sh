#!/bin/bash
# sweep the scheduling thresholds of the persistent pipeline (C4, 256 spp, one step)
run() { env "$@" python bench.py --steps 2 --warmup 1 --spp 256 --pipeline persistent --e2e-steps 0 --cpu-spp 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', round(d['value'],1))"; }
run RT_PS_WORK=24
for w in 16 20 28 32; do run RT_PS_WORK=$w; done
for s in 2 4 10 16; do run RT_PS_STALL=$s; done
for l in 1 2 8 12; do run RT_PS_LEAVE=$l; done
for d in 4 12 16 20; do run RT_PS_DESCEND=$d; done
