#!/bin/bash
# round-2 A/B (run under gpurun): C4 throughput of the persistent pipeline per BVH layout / stack placement
run() { env "$@" python bench.py --steps 3 --warmup 2 --spp 500 --pipeline ${PIPELINE:-persistent} --e2e-steps 0 --cpu-spp 0 2>gpurun_out/ab_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', round(d['value'],1), 'rays/path', round(d['rays_per_path'],3))"; }
for e in "$@"; do run $e; done
