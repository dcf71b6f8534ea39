#!/bin/bash
# short multi-GPU visit (gpurun --gpus N is charged N x its time): the sharded-render tests and the strong-scaling bench at N only
set -u
N=${1:-8}; TAG=${2:-r2s2}
timeout 600 python -m pytest tests/test_cli_and_multi.py -m gpu -x -q > gpurun_out/pytest_multi_$TAG.log 2>&1; echo pytest_rc=$?; tail -3 gpurun_out/pytest_multi_$TAG.log
out=gpurun_out/scale_${TAG}_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 3 --warmup 3 > $out 2> gpurun_out/scale_${TAG}_n$N.err
echo "C4 strong n=$N rc=$?"; tail -1 $out | head -c 1500
