#!/bin/bash
run() { env "$@" python bench.py --steps 3 --warmup 2 --spp 500 --pipeline persistent --e2e-steps 0 --cpu-spp 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', round(d['value'],1))"; }
run X=0
for e in "$@"; do run $e; done
