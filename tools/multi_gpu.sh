#!/bin/bash
# multi-GPU visit: sharded-render test, torchrun bench at every N up to the box size, C5 strong scaling through the CLI
set -u
N=${1:-2}; TAG=${2:-r1}
timeout 600 python -m pytest tests/test_cli_and_multi.py -m gpu -x -q > gpurun_out/pytest_multi_$TAG.log 2>&1; echo pytest_rc=$?; tail -3 gpurun_out/pytest_multi_$TAG.log
for n in 1 2 4 8; do
  [ $n -gt $N ] && break
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --cpu-spp 0 > gpurun_out/scale_${TAG}_n$n.json 2> gpurun_out/scale_${TAG}_n$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/scale_${TAG}_n$n.json 2> gpurun_out/scale_${TAG}_n$n.err
  fi
  echo "bench n=$n rc=$? $(python -c "import json,sys; d=json.loads(open('gpurun_out/scale_${TAG}_n$n.json').read().strip().splitlines()[-1]); print(round(d['value'],1), 'Mpaths/s e2e', round(d['e2e']['value'],1))" 2>&1 | tail -1)"
done
SPP=${C5_SPP:-4096}
for n in 1 2 4 8; do
  [ $n -gt $N ] && break
  echo "C5 strong n=$n $(./mu-lambda-raytracer_b200/rt_main --world=final_scene --seed=42 --aspect_ratio=1:1 --image_width=3840 --samples_per_pixel=$SPP --gpus $n --stats 2>&1 >/dev/null | tail -1)"
done
