#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "agree or small_pool or persistent or closest" > gpurun_out/pytest_quick.log 2>&1; echo pytest_rc=$?; tail -2 gpurun_out/pytest_quick.log
RT_PS_STATS=1 python bench.py --steps 1 --warmup 1 --spp 64 --pipeline persistent --e2e-steps 0 --cpu-spp 0 2>&1 | grep "persist stats" | tail -1
run() { env "$@" python bench.py --steps 3 --warmup 2 --spp 500 --pipeline persistent --e2e-steps 0 --cpu-spp 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', round(d['value'],1))"; }
run X=0
for e in "$@"; do run $e; done
