#!/bin/bash
# quick A/B on one GPU: pipeline agreement tests, then C4 throughput of the persistent pipeline under the given env settings
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "agree or small_pool or persistent" > gpurun_out/pytest_quick.log 2>&1; echo pytest_rc=$?; tail -2 gpurun_out/pytest_quick.log
run() { env "$@" python bench.py --steps 3 --warmup 2 --spp 500 --pipeline persistent --e2e-steps 0 --cpu-spp 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', round(d['value'],1))"; }
run X=0
for l in 1 2 3 4 6; do run RT_BVH_MAX_LEAF=$l; done
