#!/bin/bash
run() { env "$@" python bench.py --steps 2 --warmup 1 --spp 256 --pipeline persistent --e2e-steps 0 --cpu-spp 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', round(d['value'],1))"; }
for b in 5 6 7 8; do run RT_PS_MINB=$b; done
